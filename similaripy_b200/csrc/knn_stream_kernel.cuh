// knn_stream_kernel.cuh -- second-generation hot kernel (sm_100a): the expansion never stops for the drain.
//
// Replaces s_plus::compute_similarities_parallel<int,float> (reference similaripy/cython_code/s_plus.h:265-453)
// like knn_flat_kernel (knn_kernel.cuh), whose similarity arithmetic, key order and selection rules it shares.
// What changes is the schedule inside the CTA (one persistent CTA of 1024 threads per SM):
//
//   warps 5..31  EXPAND   stream B as 16-byte chunks of two (column, value) pairs with cp.async (LDGSTS) into a
//                         lane-private shared-memory ring -- the next batch is in flight while the current one is added
//                         to the panel with red.shared.add.f32 (s_plus.h:358-403 / 418-438).  The chunks of a pass are
//                         numbered through by an exclusive scan of the segments' chunk counts and cut into 27 equal
//                         ranges, one per warp, whatever the segment lengths: every lane of every warp has work, and
//                         a long B row is shared by several warps;
//   warp  4      STAGE    claims target rows from the atomic queue (`omp for schedule(dynamic)`, s_plus.h:337) and
//                         prepares the NEXT pass while the current one is expanded: per entry of the target row the
//                         chunk range of its B-row segment inside the panel (read coalesced from the per-call
//                         `aexp` table -- no dependent split-point lookup on the critical path), the value of the
//                         entry, and the scan;
//   warps 4..31  SNAPSHOT when a panel is complete: LDS.128 -> STS.128 (reset to "untouched") -> tcgen05.st: the panel
//                         moves to TENSOR MEMORY (256 KB per SM, idle in a kernel without MMAs) and the expansion of
//                         the next panel starts at once;
//   warps 0..3   DRAIN    read the snapshot back with tcgen05.ld (each warp its lane quarter), apply the coarse
//                         per-128-column bound, the per-slot pre-filter, computeSimilarity (s_plus.h:129-156), the
//                         threshold (s_plus.h:206) and keep the best k (s_plus.h:45-59, 193-215), concurrently with
//                         the expansion.  TMEM is read-only for the drain, so an overflowing candidate buffer just
//                         selects and resumes.
//   handshake: two mbarriers (snapshot full / snapshot free) and a two-entry message ring; named barriers inside the
//   roles.  Every wait is bounded (trap + error flag) so that a protocol bug cannot hang the device.
//
// TMEM mapping of a panel: column c, quad q = c / 4, tile T = q / 128 (512 columns): lane = q % 128, TMEM column
// = 4 T + c % 4.  A warp can only touch lane quarter (warp id % 4).
#pragma once
#include "knn_kernel.cuh"

namespace spy {

struct KnnStreamDev {
    KnnDev q;                // shared fields (targets, A, vectors, scalars, k, cap, selectors, outputs, work counter)
    const long long *toff;   // [n_targets + 1]: first compact entry of target row i (exclusive scan of the row lengths)
    long long E;             // toff[n_targets]
    const uint2 *aexp;       // [n_panels][E]: (first chunk, end chunk) of the entry's B-row segment inside the panel
    const uint4 *chunks;     // B as 16-byte chunks of two (column, value) pairs; every row padded to whole chunks
    const float *ymin_t, *ymin_c, *ymin_d;  // minima of Yt / Yc / Yd over every 128 consecutive columns (NULL: unused)
    int *err;                // set to non-zero before a trap (bounded waits)
};

constexpr int KS_NT = 1024;
constexpr int KS_D_WARPS = 4;                  // warps 0..3: drain
constexpr int KS_S_WARP = 4;                   // warp 4: stage
constexpr int KS_A_WARPS = 27;                 // warps 5..31: expand
constexpr int KS_X_THREADS = (KS_A_WARPS + 1) * 32;  // warps 4..31: the threads of the expansion-side barriers
constexpr int KS_U = 2;                        // 16-byte chunks per lane per batch (ring: 32 bytes per lane)
constexpr int KS_CH = 1024;                    // entries of a target row per pass
constexpr int KS_QCAP = 64;                    // quads per drain warp waiting for their per-slot test
constexpr int KS_FLAG_PANEL_END = 1, KS_FLAG_ROW_END = 2, KS_FLAG_STOP = 4;

struct KsPass {   // what the stage warp hands to the expansion warps (shared memory, double buffered)
    int n;        // entries staged
    unsigned total;  // chunks of the pass
    int pn, flags, t, i_out;
};
struct KsMsg {    // what the expansion side hands to the drain with every snapshot (shared memory, double buffered)
    int t, i_out, pn, flags, landed;
};

__host__ __device__ constexpr size_t ks_ring_bytes() { return (size_t)KS_A_WARPS * 32 * KS_U * 16; }
__host__ __device__ constexpr size_t ks_stage_bytes() { return (size_t)2 * ((KS_CH + 32) * 4 + KS_CH * 4 + KS_CH * 4); }
__host__ __device__ constexpr size_t ks_queue_bytes() { return (size_t)KS_D_WARPS * KS_QCAP * (16 + 4); }
__host__ __device__ constexpr size_t ks_fixed_bytes(int cap) { return ks_ring_bytes() + ks_stage_bytes() + ks_queue_bytes() + (size_t)cap * 8; }

// ---- PTX helpers --------------------------------------------------------------------------------------------
__device__ __forceinline__ void ks_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ bool ks_bar_or(int id, int count, bool pred) {
    unsigned out;
    asm volatile("{ .reg .pred p, q; setp.ne.u32 q, %3, 0; bar.red.or.pred p, %1, %2, q; selp.u32 %0, 1, 0, p; }"
                 : "=r"(out) : "r"(id), "r"(count), "r"((unsigned)pred) : "memory");
    return out != 0;
}
__device__ __forceinline__ void ks_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ks_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool ks_mbar_try(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait (about two seconds): a protocol bug must end in an error, not in a hung device
__device__ __noinline__ void ks_timeout(int *err, int code) {
    if (err) atomicExch(err, code);
    __threadfence_system();
    __trap();
}
__device__ __forceinline__ void ks_mbar_wait(unsigned bar, unsigned parity, int *err, int code) {
    if (ks_mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!ks_mbar_try(bar, parity))
        if (clock64() - t0 > 4000000000LL) ks_timeout(err, code);
}
__device__ __forceinline__ void ks_cp_async16(unsigned dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ks_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ks_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 ks_lds128u(unsigned a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void ks_tmem_st4(unsigned taddr, float4 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v.x)),
                 "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
}
__device__ __forceinline__ float4 ks_tmem_ld4(unsigned taddr) {
    unsigned r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
}
__device__ __forceinline__ void ks_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ks_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- drain side: sort / select with the 128 threads of the drain warps ---------------------------------------
constexpr int KS_DT = KS_D_WARPS * 32;
__device__ __forceinline__ void ks_dsync() { ks_bar_sync(4, KS_DT); }

// Bitonic sort (descending) of cand[0, S), S a power of two; the drain warps only.
__device__ void ks_bitonic_desc(u64 *cand, int S, int dtid) {
    for (int size = 2; size <= S; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = dtid; i < (S >> 1); i += KS_DT) {
                const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0) || (size == S);
                const u64 a = cand[lo], b = cand[hi];
                if ((a < b) == desc) { cand[lo] = b; cand[hi] = a; }
            }
            ks_dsync();
        }
    }
}
// Exact selection among the evaluated keys cand[0, n) (0 = dead): leaves the best m = min(k, live) sorted best-first
// in cand[0, m); returns m and, when m == k, sets tau to the k-th key.  (The reference's heap, s_plus.h:45-59.)
__device__ int ks_select(u64 *cand, int n, int k, u64 &tau, int *s_live, int dtid) {
    int S = 32;
    while (S < n) S <<= 1;
    for (int i = n + dtid; i < S; i += KS_DT) cand[i] = 0ull;
    if (dtid == 0) *s_live = 0;
    ks_dsync();
    ks_bitonic_desc(cand, S, dtid);
    for (int i = dtid; i < S; i += KS_DT)  // live keys are a prefix: find its end
        if (cand[i] != 0ull && (i == S - 1 || cand[i + 1] == 0ull)) *s_live = i + 1;
    ks_dsync();
    const int m = min(*s_live, k);
    if (m == k) tau = cand[k - 1];
    ks_dsync();
    return m;
}

// The kernel.  KIND selects the drain's pre-filter like in knn_flat_kernel (KIND_RAW / _T / _C / _D / _GEN).
template <int KIND>
__global__ void __launch_bounds__(KS_NT, 1)
knn_stream_kernel(const __grid_constant__ KnnStreamDev p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const KnnDev &q = p.q;
    float *acc = reinterpret_cast<float *>(smem_raw);
    unsigned char *ptr = smem_raw + (size_t)q.W * sizeof(float);
    const unsigned ring32 = (unsigned)__cvta_generic_to_shared(ptr);
    ptr += ks_ring_bytes();
    unsigned *stP[2], *stC[2];
    float *stV[2];
    for (int b = 0; b < 2; b++) {
        stP[b] = reinterpret_cast<unsigned *>(ptr); ptr += (KS_CH + 32) * 4;
        stC[b] = reinterpret_cast<unsigned *>(ptr); ptr += KS_CH * 4;
        stV[b] = reinterpret_cast<float *>(ptr); ptr += KS_CH * 4;
    }
    float4 *qx_all = reinterpret_cast<float4 *>(ptr); ptr += (size_t)KS_D_WARPS * KS_QCAP * 16;
    int *qc_all = reinterpret_cast<int *>(ptr); ptr += (size_t)KS_D_WARPS * KS_QCAP * 4;
    u64 *cand = reinterpret_cast<u64 *>(ptr);

    __shared__ __align__(8) unsigned long long s_full, s_empty;
    __shared__ KsPass s_pass[2];
    __shared__ KsMsg s_msg[2];
    __shared__ unsigned s_tmem;
    __shared__ int s_cnt, s_overflow, s_live;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float sentinel = __uint_as_float(kSentinelBits);
    const float4 sentinel4 = make_float4(sentinel, sentinel, sentinel, sentinel);
    const unsigned acc32 = (unsigned)__cvta_generic_to_shared(acc);
    const unsigned full32 = (unsigned)__cvta_generic_to_shared(&s_full), empty32 = (unsigned)__cvta_generic_to_shared(&s_empty);
    const int nT = q.W >> 9;  // tiles of 512 columns (W is a multiple of 512)

    for (int i = tid * 4; i < q.W; i += KS_NT * 4) *reinterpret_cast<float4 *>(acc + i) = sentinel4;
    if (tid == 0) {
        ks_mbar_init(full32, 1);
        ks_mbar_init(empty32, KS_D_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_cnt = 0; s_overflow = 0; s_live = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ks_tc_fence_before();
    __syncthreads();
    ks_tc_fence_after();
    const unsigned tmem_q = s_tmem + (((unsigned)(warp & 3) * 32u) << 16);  // this warp's lane quarter

    if (warp >= KS_D_WARPS) {
        // =====================================================================================================
        // expansion side: warp 4 stages, warps 5..31 expand, all 28 take the snapshots
        // =====================================================================================================
        const bool is_stage = warp == KS_S_WARP;
        const int wa = warp - (KS_S_WARP + 1);  // 0..26 for the expansion warps
        // ---- stage warp state (uniform registers) ----
        bool s_have = false;
        int s_i_out = 0, s_t = 0, s_a0 = 0, s_len = 0, s_pn = 0, s_c0 = 0;
        long long s_tq0 = 0;
        auto stage_next = [&](int buf) {  // prepare the next pass into buffer `buf`
            KsPass *d = &s_pass[buf];
            if (!s_have) {
                int slot = 0;
                if (lane == 0) slot = atomicAdd(q.work_counter, 1);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot >= q.n_targets) {
                    if (lane == 0) { d->n = 0; d->total = 0u; d->pn = 0; d->flags = KS_FLAG_STOP; d->t = 0; d->i_out = 0; }
                    return;
                }
                s_i_out = q.row_order ? __ldg(q.row_order + slot) : slot;
                s_t = __ldg(q.targets + s_i_out);
                s_a0 = __ldg(q.a_indptr + s_t);
                s_len = __ldg(q.a_indptr + s_t + 1) - s_a0;
                s_tq0 = __ldg(p.toff + s_i_out);
                s_pn = 0; s_c0 = 0; s_have = true;
                if (s_len <= 0) {  // empty row: one message so that the drain writes its (empty) output
                    if (lane == 0) { d->n = 0; d->total = 0u; d->pn = q.n_panels - 1; d->flags = KS_FLAG_PANEL_END | KS_FLAG_ROW_END; d->t = s_t; d->i_out = s_i_out; }
                    s_have = false;
                    return;
                }
            }
            const int n = min(KS_CH, s_len - s_c0);
            const uint2 *src = p.aexp + ((long long)s_pn * p.E + s_tq0 + s_c0);
            const float *vsrc = q.a_data + s_a0 + s_c0;
            unsigned *sP = stP[buf], *sC = stC[buf];
            float *sV = stV[buf];
            unsigned carry = 0u;
            for (int r0 = 0; r0 < n; r0 += 256) {
                uint2 se[8];
                float vv[8];
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const int i = r0 + r * 32 + lane;
                    se[r] = make_uint2(0u, 0u); vv[r] = 0.f;
                    if (i < n) { se[r] = __ldg(src + i); vv[r] = __ldg(vsrc + i); }
                }
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const int i = r0 + r * 32 + lane;
                    if (r0 + r * 32 < n) {  // uniform
                        const unsigned c = se[r].y - se[r].x;
                        unsigned inc = c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o) inc += v;
                        }
                        if (i < n) { sP[i] = carry + inc - c; sC[i] = se[r].x; sV[i] = vv[r]; }
                        carry += __shfl_sync(0xffffffffu, inc, 31);
                    }
                }
            }
            int flags = 0;
            if (s_c0 + n == s_len) {
                flags |= KS_FLAG_PANEL_END;
                if (s_pn == q.n_panels - 1) flags |= KS_FLAG_ROW_END;
            }
            if (lane == 0) { sP[n] = carry; d->n = n; d->total = carry; d->pn = s_pn; d->flags = flags; d->t = s_t; d->i_out = s_i_out; }
            s_c0 += n;
            if (s_c0 == s_len) { s_c0 = 0; s_pn++; if (s_pn == q.n_panels) s_have = false; }
        };

        // ---- expansion warp state ----
        const unsigned slot32 = ring32 + (unsigned)((max(wa, 0) * 32 * KS_U + lane) * 16);  // chunk r of a batch: + r * 512
        unsigned f = 0u, fb = 0u, F1 = 0u;  // next chunk to issue, end of the sub-batch, end of the warp's range (uniform)
        int j0 = 0, n_cur = 0;
        unsigned Pe = 0u, Pi = 0u, C0 = 0u;  // lane l: segment j0 + l of the pass: first chunk number, end, first chunk in `chunks`
        float V = 0.f;
        const unsigned *cP = nullptr, *cC = nullptr;
        const float *cV = nullptr;
        float vp[KS_U];
        unsigned lp = 0u;
#pragma unroll
        for (int r = 0; r < KS_U; r++) vp[r] = 0.f;
        auto issue = [&]() {  // one batch of KS_U chunks per lane into the ring
            if (f >= fb) {    // next 32 segments (uniform)
                const int idx = j0 + lane;
                Pe = cP[min(idx, n_cur)];
                Pi = cP[min(idx + 1, n_cur)];
                C0 = cC[min(idx, n_cur - 1)];
                V = cV[min(idx, n_cur - 1)];
                fb = min(F1, __shfl_sync(0xffffffffu, Pi, 31));
                j0 += 32;
            }
            lp = 0u;
#pragma unroll
            for (int r = 0; r < KS_U; r++) {
                const unsigned fr = f + (unsigned)(r * 32 + lane);
                const bool live = fr < fb;
                int j = 0;  // number of segments of the sub-batch that end at or before chunk fr
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const unsigned t = __shfl_sync(0xffffffffu, Pi, j + step - 1);
                    if (t <= fr) j += step;
                }
                const unsigned pe = __shfl_sync(0xffffffffu, Pe, j), c0 = __shfl_sync(0xffffffffu, C0, j);
                vp[r] = __shfl_sync(0xffffffffu, V, j);
                if (live) {
                    ks_cp_async16(slot32 + (unsigned)r * 512u, p.chunks + (c0 + (fr - pe)));
                    lp |= 1u << r;
                }
            }
            ks_cp_commit();
            f = min(f + 32u * KS_U, fb);
        };

        // ---- snapshot bookkeeping (uniform over the 28 warps) ----
        unsigned seq = 0u;  // snapshots / messages posted so far
        bool pending = false, panel_landed = false;
        int m_t = 0, m_i_out = 0, m_pn = 0, m_flags = 0, m_landed = 0;
        auto post = [&](int t, int i_out, int pn, int flags, int landed) {  // one thread: message + "snapshot full"
            KsMsg *m = &s_msg[seq & 1u];
            m->t = t; m->i_out = i_out; m->pn = pn; m->flags = flags; m->landed = landed;
            ks_mbar_arrive(full32);
        };
        auto snapshot = [&]() {
            // the drain must have released the previous snapshot (and read its message)
            if (lane == 0) ks_mbar_wait(empty32, (seq & 1u) ^ 1u, p.err, 2);
            __syncwarp();
            ks_tc_fence_after();
            if (m_landed) {
                const int base = m_pn * q.W;
                if (q.filter_mode == SPY_SEL_MATRIX) {  // erase the row's filtered columns first (s_plus.h:159-172)
                    const int width = min(q.W, q.n_cols - base);
                    const int fs = __ldg(q.f_indptr + m_t), fe = __ldg(q.f_indptr + m_t + 1);
                    for (int x = fs + (tid - KS_D_WARPS * 32); x < fe; x += KS_X_THREADS) {
                        const int c = __ldg(q.f_indices + x) - base;
                        if (c >= 0 && c < width) acc[c] = sentinel;
                    }
                    ks_bar_sync(3, KS_X_THREADS);
                }
                const int sub = (warp - KS_D_WARPS) >> 2;  // 0..6: the warps of one lane quarter
                for (int T = sub; T < nT; T += 7) {
                    const unsigned a = acc32 + (unsigned)(512 * T + 128 * (warp & 3) + 4 * lane) * 4u;
                    const float4 x = lds128(a);
                    sts128(a, sentinel4);
                    ks_tmem_st4(tmem_q + (unsigned)(4 * T), x);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            ks_tc_fence_before();
            ks_bar_sync(2, KS_X_THREADS);  // every slot of the panel is in TMEM and reset
            if (tid == KS_D_WARPS * 32) post(m_t, m_i_out, m_pn, m_flags, m_landed);
            seq++;
        };

        if (is_stage) stage_next(0);
        ks_bar_sync(1, KS_X_THREADS);
        for (int pass = 0;; pass++) {
            const int buf = pass & 1;
            const KsPass d = s_pass[buf];
            const bool stop = (d.flags & KS_FLAG_STOP) != 0;
            bool have = false;
            unsigned accb32 = 0u, base = 0u;
            if (!is_stage && !stop) {  // this warp's chunk range of the pass and its first batch
                n_cur = d.n; cP = stP[buf]; cC = stC[buf]; cV = stV[buf];
                base = (unsigned)d.pn * (unsigned)q.W;
                accb32 = acc32 - base * 4u;
                const unsigned F0 = (unsigned)(((unsigned long long)d.total * (unsigned)wa) / KS_A_WARPS);
                F1 = (unsigned)(((unsigned long long)d.total * (unsigned)(wa + 1)) / KS_A_WARPS);
                f = F0; fb = F0;
                if (F0 < F1) {
                    // j0 = the last segment that starts at or before chunk F0 (two rounds of 32 probes)
                    int idx = lane * 32;
                    unsigned v = idx <= n_cur ? cP[idx] : 0xffffffffu;
                    const int blk = __popc(__ballot_sync(0xffffffffu, v <= F0)) - 1;
                    idx = blk * 32 + lane;
                    v = idx <= n_cur ? cP[idx] : 0xffffffffu;
                    j0 = blk * 32 + __popc(__ballot_sync(0xffffffffu, v <= F0)) - 1;
                    issue();
                    have = true;
                }
            }
            if (pending) snapshot();
            if (stop) break;
            bool any = false;
            if (is_stage) stage_next(buf ^ 1);
            else {
                any = have;
                while (have) {
                    ks_cp_wait_all();
                    uint4 pr[KS_U];
                    float vc[KS_U];
#pragma unroll
                    for (int r = 0; r < KS_U; r++) { pr[r] = ks_lds128u(slot32 + (unsigned)r * 512u); vc[r] = vp[r]; }
                    const unsigned lc = lp;
                    const bool more = f < F1;
                    if (more) issue();
#pragma unroll
                    for (int r = 0; r < KS_U; r++) {
                        if (lc & (1u << r)) {
                            const unsigned d0 = pr[r].x - base, d1 = pr[r].z - base;
                            if (d0 < (unsigned)q.W) smem_add_f32(acc32 + d0 * 4u, __fmul_rn(__uint_as_float(pr[r].y), vc[r]));
                            if (d1 < (unsigned)q.W) smem_add_f32(acc32 + d1 * 4u, __fmul_rn(__uint_as_float(pr[r].w), vc[r]));
                        }
                    }
                    have = more;
                }
            }
            (void)accb32;
            const bool landed = ks_bar_or(1, KS_X_THREADS, any);  // all adds of the pass have landed; the next pass is staged
            panel_landed |= landed;
            pending = false;
            if (d.flags & KS_FLAG_PANEL_END) {
                if (panel_landed || (d.flags & KS_FLAG_ROW_END)) {
                    pending = true;
                    m_t = d.t; m_i_out = d.i_out; m_pn = d.pn; m_flags = d.flags; m_landed = panel_landed ? 1 : 0;
                }
                panel_landed = false;
            }
        }
        // no more rows: tell the drain
        if (tid == KS_D_WARPS * 32) {
            ks_mbar_wait(empty32, (seq & 1u) ^ 1u, p.err, 3);
            post(0, 0, 0, KS_FLAG_STOP, 0);
        }
    } else {
        // =====================================================================================================
        // drain side: warps 0..3, warp w reads lane quarter w of the snapshot
        // =====================================================================================================
        const int dtid = tid;  // 0..127
        float4 *qx = qx_all + warp * KS_QCAP;
        int *qc = qc_all + warp * KS_QCAP;
        const bool filter = !q.exact_only;
        const bool useT = filter && (KIND == KIND_T || (KIND == KIND_GEN && q.l1 != 0.f));
        const bool useC = filter && (KIND == KIND_C || (KIND == KIND_GEN && q.l2 != 0.f));
        const bool useD = filter && (KIND == KIND_D || (KIND == KIND_GEN && q.l3 != 0.f));
        bool row_open = false;
        SimRow sr = {0.f, 0.f, 0.f};
        FastRow fr = {0.f, 0.f, 0.f, 0.f, 0.f};
        u64 tau = 0ull;
        float lo = reject_bound(q, tau);
        int n_eval = 0;  // cand[0, n_eval) evaluated keys, cand[n_eval, s_cnt) raw candidates
        // exact values of the raw candidates (computeSimilarity, s_plus.h:129-156; threshold, s_plus.h:206)
        auto evaluate = [&](int cnt) {
            for (int i = n_eval + dtid; i < cnt; i += KS_DT) {
                const u64 raw = cand[i];
                const int col = (int)(unsigned)(raw & 0xffffffffull);
                u64 key = 0ull;
                if (col >= 0) {  // (column -1: filler of a reservation that did not fit)
                    const float val = similarity_value(q, sr, __uint_as_float((unsigned)(raw >> 32)),
                                                       q.l1 != 0.f ? __ldg(q.Yt + col) : 0.f, q.l2 != 0.f ? __ldg(q.Yc + col) : 0.f,
                                                       q.l3 != 0.f ? __ldg(q.Yd + col) : 0.f);
                    if (val >= q.thr) key = make_key(val, col);
                }
                cand[i] = (key > tau) ? key : 0ull;
            }
            n_eval = cnt;
        };
        auto select_now = [&]() {  // evaluate what is buffered, keep the best k, raise tau
            ks_dsync();
            const int cnt = min(*reinterpret_cast<volatile int *>(&s_cnt), q.cap);
            evaluate(cnt);
            const int m = ks_select(cand, cnt, q.k, tau, &s_live, dtid);
            lo = reject_bound(q, tau);
            n_eval = m;
            if (dtid == 0) { s_cnt = m; s_overflow = 0; }
            ks_dsync();
        };

        for (unsigned seq = 0u;; seq++) {
            if (lane == 0) ks_mbar_wait(full32, seq & 1u, p.err, 1);
            __syncwarp();
            const KsMsg m = s_msg[seq & 1u];
            if (m.flags & KS_FLAG_STOP) break;
            ks_tc_fence_after();
            if (!row_open) {
                row_open = true;
                sr.Xt = (q.l1 != 0.f) ? __ldg(q.Xt + m.t) : 0.f;
                sr.Xc = (q.l2 != 0.f) ? __ldg(q.Xc + m.t) : 0.f;
                sr.Xd = (q.l3 != 0.f) ? __ldg(q.Xd + m.t) : 0.f;
                fr.A0 = q.stab + q.l1 * q.t1 * sr.Xt;
                fr.cT = q.l1 * q.t2;
                fr.cX = q.l1 * (1.f - q.t1 - q.t2);
                fr.cC = q.l2 * sr.Xc;
                fr.cD = q.l3 * sr.Xd;
                tau = 0ull;
                lo = reject_bound(q, tau);
                n_eval = 0;  // (s_cnt was reset when the previous row was written)
            }
            if (m.landed) {
                const int base = m.pn * q.W;
                int Tcur = 0, qn = 0;  // next tile of this warp's sweep; quads waiting in the warp's queue
                for (;;) {  // sweep; leaves the loop when the warp's share is done; re-entered after an overflow
                    bool overflow = false;
                    // coarse bound of a tile (per 128-column block of this warp's lane quarter): a slot can only enter the
                    // result if  x >= lo * den  and  den >= Dmin + cX * x  with Dmin from the block minima of Y, i.e.
                    // x * (1 - lo * cX) >= lo * Dmin.  Lane l holds the bounds of tiles l, l + 32, l + 64, l + 96.
                    const float g = 1.f - lo * fr.cX;
                    const bool coarse = filter && lo > 0.f && g > 0.f && fr.cT >= 0.f && fr.cC >= 0.f && fr.cD >= 0.f &&
                                        (!useT || p.ymin_t != nullptr) && (!useC || p.ymin_c != nullptr) && (!useD || p.ymin_d != nullptr);
                    float cb[4] = {0.f, 0.f, 0.f, 0.f};
                    if (coarse) {
#pragma unroll
                        for (int ri = 0; ri < 4; ri++) {
                            const int T = lane + 32 * ri;
                            const int blk = (base >> 7) + 4 * T + warp;
                            if (T < nT && blk < ((q.n_cols + 127) >> 7)) {
                                float dmin = fr.A0;
                                if (useT) dmin = fmaf(fr.cT, __ldg(p.ymin_t + blk), dmin);
                                if (useC) dmin = fmaf(fr.cC, __ldg(p.ymin_c + blk), dmin);
                                if (useD) dmin = fmaf(fr.cD, __ldg(p.ymin_d + blk), dmin);
                                cb[ri] = (KIND == KIND_RAW) ? lo : (dmin > 0.f ? lo * dmin / g : 0.f);
                            }
                        }
                    }
                    while (Tcur < nT || qn > 0) {
                        if (qn >= 32 || Tcur >= nT) {
                            // ---- per-slot test of up to 32 queued quads, all lanes busy: one L2 round trip for the batch ----
                            const int nb = min(qn, 32);
                            const int e = qn - nb + lane;  // take from the end of the queue
                            unsigned sm = 0u;
                            float4 x = sentinel4;
                            int col0 = 0;
                            if (lane < nb) {
                                x = qx[e]; col0 = qc[e];
                                float4 yt = sentinel4, yc = sentinel4, yd = sentinel4;
                                if (col0 + 3 < q.n_cols) {
                                    if (useT) yt = __ldg(reinterpret_cast<const float4 *>(q.Yt + col0));
                                    if (useC) yc = __ldg(reinterpret_cast<const float4 *>(q.Yc + col0));
                                    if (useD) yd = __ldg(reinterpret_cast<const float4 *>(q.Yd + col0));
                                } else {
                                    if (useT) yt = load_y4(q.Yt, col0, q.n_cols);
                                    if (useC) yc = load_y4(q.Yc, col0, q.n_cols);
                                    if (useD) yd = load_y4(q.Yd, col0, q.n_cols);
                                }
                                const float lc = lo * (KIND == KIND_D ? fr.cD : fr.cC), la = lo * fr.A0;
                                sm = survivor_mask<KIND>(q, fr, filter, lo, lc, la, x, yt, yc, yd);
                            }
                            const int c = __popc(sm);
                            int inc = c;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                                if (lane >= o) inc += v;
                            }
                            const int total = __shfl_sync(0xffffffffu, inc, 31);
                            if (total > 0) {
                                int pos = 0;
                                if (lane == 31) pos = atomicAdd(&s_cnt, total);
                                pos = __shfl_sync(0xffffffffu, pos, 31);
                                if (pos + total > q.cap) {  // does not fit: dead fillers, select, come back
                                    for (int i = pos + lane; i < min(pos + total, q.cap); i += 32) cand[i] = 0xffffffffull;
                                    overflow = true;
                                    break;
                                }
                                int w = pos + inc - c;
                                const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                                for (int r = 0; r < 4; r++)
                                    if (sm & (1u << r)) cand[w++] = make_raw(xs[r], col0 + r);
                            }
                            qn -= nb;
                            continue;
                        }
                        // ---- coarse test of the next tile: quads that cannot be rejected as a whole join the queue ----
                        const float4 x = ks_tmem_ld4(tmem_q + (unsigned)(4 * Tcur));
                        bool pass;
                        const bool touched = (__float_as_uint(x.x) != kSentinelBits) | (__float_as_uint(x.y) != kSentinelBits) |
                                             (__float_as_uint(x.z) != kSentinelBits) | (__float_as_uint(x.w) != kSentinelBits);
                        if (coarse) {
                            const int ri = Tcur >> 5;
                            const float mine = ri == 0 ? cb[0] : ri == 1 ? cb[1] : ri == 2 ? cb[2] : cb[3];
                            const float bound = __shfl_sync(0xffffffffu, mine, Tcur & 31);
                            if (KIND == KIND_RAW || KIND == KIND_C || KIND == KIND_D)  // den >= 0: a negative dot product never passes lo > 0
                                pass = !((x.x < bound) & (x.y < bound) & (x.z < bound) & (x.w < bound));
                            else  // a negative dot product over a negative denominator is left to the per-slot test
                                pass = !((x.x < bound) & (x.y < bound) & (x.z < bound) & (x.w < bound)) | (x.x < 0.f) | (x.y < 0.f) |
                                       (x.z < 0.f) | (x.w < 0.f);
                            pass &= touched;
                        } else pass = touched;
                        const unsigned bal = __ballot_sync(0xffffffffu, pass);
                        if (pass) {
                            const int e = qn + __popc(bal & ((1u << lane) - 1u));
                            qx[e] = x;
                            qc[e] = base + 512 * Tcur + 128 * warp + 4 * lane;
                        }
                        qn += __popc(bal);
                        Tcur++;
                        __syncwarp();
                    }
                    if (overflow && lane == 0) s_overflow = 1;
                    ks_dsync();  // every drain warp is done or stopped at a full buffer
                    const bool again = *reinterpret_cast<volatile int *>(&s_overflow) != 0;
                    if (!again) break;
                    select_now();
                }
            }
            // the snapshot has been read: hand TMEM back (and with it the message slot)
            ks_tc_fence_before();
            __syncwarp();
            if (lane == 0) ks_mbar_arrive(empty32);
            // evaluate what this panel added; tighten the bound while the buffer is reasonably full
            ks_dsync();
            {
                const int cnt = min(*reinterpret_cast<volatile int *>(&s_cnt), q.cap);
                if (cnt > q.cap / 2 && !(m.flags & KS_FLAG_ROW_END)) select_now();
                else { evaluate(cnt); ks_dsync(); }
            }
            if (m.flags & KS_FLAG_ROW_END) {
                // ---- final selection and slab write (s_plus.h:443-450) ----
                const int n_out = ks_select(cand, n_eval, q.k, tau, &s_live, dtid);
                const size_t o = (size_t)m.i_out * (size_t)q.k;
                for (int j = dtid; j < q.k; j += KS_DT) {
                    int col = 0, row = 0;
                    float val = 0.f;
                    if (j < n_out) {
                        const u64 key = cand[j];
                        col = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
                        val = unordered_bits((unsigned)(key >> 32));
                        row = m.t;
                    }
                    q.out_cols[o + j] = col;
                    q.out_vals[o + j] = val;
                    if (q.out_rows) q.out_rows[o + j] = row;
                }
                if (dtid == 0) {
                    if (q.out_counts) q.out_counts[m.i_out] = n_out;
                    s_cnt = 0;
                }
                row_open = false;
                ks_dsync();
            }
        }
    }
    ks_tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_tmem) : "memory");
}

typedef void (*knn_stream_kernel_t)(const KnnStreamDev);

}  // namespace spy

// ---- host side (knn_stream.cu), used by the planner / launcher in knn_kernel.cu ------------------------------
namespace spy {
struct StreamPlan {
    int W, n_panels, cap;
    size_t smem_bytes;
};
// false when the stream engine does not cover the configuration (k too large for the shared-memory budget)
bool stream_plan(int k, int n_cols, int panel_width, int max_smem_optin, StreamPlan &sp);
int64_t stream_scratch_bytes(int n_cols);
int stream_launch(const spy_knn_args &a, const StreamPlan &sp, int kind, int exact_only, int grid, void *scratch,
                  int64_t scratch_bytes, cudaStream_t st);
}  // namespace spy
