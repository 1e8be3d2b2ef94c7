// knn_stream_kernel.cuh -- second-generation hot kernel (sm_100a): the expansion never stops for the drain.
//
// Replaces s_plus::compute_similarities_parallel<int,float> (reference similaripy/cython_code/s_plus.h:265-453)
// like knn_flat_kernel (knn_kernel.cuh), whose similarity arithmetic, key order and selection rules it shares.
// What changes is the schedule inside the CTA (one persistent CTA of 1024 threads per SM), D = KS_D_WARPS:
//
//   warps D..31  EXPAND   stream B as 16-byte chunks of two (column, value) pairs with cp.async (LDGSTS) into a
//                         lane-private shared-memory ring -- the next batch is in flight while the current one is added
//                         to the panel with red.shared.add.f32 (s_plus.h:358-403 / 418-438).  The chunks of a pass are
//                         numbered through by an exclusive scan of the segments' chunk counts and cut into equal
//                         ranges, one per warp, whatever the segment lengths: every lane of every warp has work, and
//                         a long B row is shared by several warps.  The same warps STAGE the next pass while the
//                         current one runs: per entry of the target row the chunk range of its B-row segment inside
//                         the panel (read coalesced from the per-call `aexp` table -- no dependent split-point lookup
//                         on the critical path), the entry's value, and a per-warp scan of the chunk counts;
//   warp  D      QUEUE    also claims target rows from the atomic queue (`omp for schedule(dynamic)`, s_plus.h:337) and
//                         resolves the pass descriptors two passes ahead;
//   warps D..31  SNAPSHOT when a panel is complete: LDS.128 -> STS.128 (reset to "untouched") -> tcgen05.st: the panel
//                         moves to TENSOR MEMORY (256 KB per SM, idle in a kernel without MMAs) and the expansion of
//                         the next panel starts at once;
//   warps 0..D-1 DRAIN    read the snapshot back with tcgen05.ld (each warp its lane quarter), apply the coarse
//                         per-128-column bound, the per-slot pre-filter, computeSimilarity (s_plus.h:129-156), the
//                         threshold (s_plus.h:206) and keep the best k (s_plus.h:45-59, 193-215), concurrently with
//                         the expansion.  TMEM is read-only for the drain, so an overflowing candidate buffer just
//                         selects and resumes.
//   handshake: two pairs of mbarriers (snapshot full / snapshot free; hand-over n uses pair n & 1, so that sparse hand-overs
//   can be double buffered in TMEM) and a two-entry message ring; named barriers inside the
//   roles.  Every wait is bounded (trap + error flag) so that a protocol bug cannot hang the device.
//
// TMEM mapping of a panel: column c, quad q = c / 4, tile T = q / 128 (512 columns): lane = q % 128, TMEM column
// = 4 T + c % 4.  A warp can only touch lane quarter (warp id % 4).
#pragma once
#include "knn_kernel.cuh"

// The kernel is compiled twice with different splits of the CTA's 32 warps between the drain and the expansion (knn_stream.cu:
// 8 + 24; knn_stream_sparse.cu: 16 + 16 for target rows with few products per panel, where the drain is the critical
// path).  Everything that depends on the split lives in an inline namespace named after the build, so that the two
// translation units do not define the same symbols.
#ifndef SPY_KS_TAG
#define SPY_KS_TAG a
#endif
#define SPY_KS_CAT2(x, y) x##y
#define SPY_KS_CAT(x, y) SPY_KS_CAT2(x, y)
#define SPY_KS_INL SPY_KS_CAT(ks_, SPY_KS_TAG)

namespace spy {
inline namespace SPY_KS_INL {

struct KnnStreamDev {
    KnnDev q;                // shared fields (targets, A, vectors, scalars, k, cap, selectors, outputs, work counter)
    const long long *toff;   // [n_targets + 1]: first compact entry of target row i (exclusive scan of the row lengths)
    long long E;             // toff[n_targets]
    const uint2 *aexp;       // [n_panels][E]: (first chunk, end chunk) of the entry's B-row segment inside the panel
    const uint4 *chunks;     // B as 16-byte chunks of two (column, value) pairs; every row padded to whole chunks
    const float *ymin_t, *ymin_c, *ymin_d;  // [n_panels][128]: minimum of Yt / Yc / Yd over the columns a TMEM lane holds of a panel
                                            // (quads lane, lane + 128, ... of the panel); NULL: unused
    int *err;                // set to non-zero before a trap (bounded waits)
};

#ifndef SPY_KS_D_WARPS
#define SPY_KS_D_WARPS 8
#endif
#ifndef SPY_KS_U
#define SPY_KS_U 2
#endif
#ifndef SPY_KS_RING
#define SPY_KS_RING 1  // 1: cp.async (LDGSTS) ring in shared memory; 0: the next batch waits in registers (LDG.128)
#endif
#ifndef SPY_KS_SPEC
#define SPY_KS_SPEC 1  // speculative bound for a row's first panel (validated; see the drain)
#endif
#ifndef SPY_KS_SEARCH
#define SPY_KS_SEARCH 1  // 1: segment of a chunk from a bit mask of segment starts (2 REDUX.OR per batch); 0: 5-step binary search by shuffles
#endif
#ifndef SPY_KS_LOCAL
#define SPY_KS_LOCAL 1   // 1: chunks hold (byte offset of the slot inside its panel, value); 0: (column, value) -- see pad_chunks_kernel
#endif
#ifndef SPY_KS_HINT
#define SPY_KS_HINT 1    // a row's first panel is swept with the bound the previous row's first panel validated (validated again)
#endif
#ifndef SPY_KS_REGSORT
#define SPY_KS_REGSORT 1  // the exact selection sorts in registers (shuffles; shared memory only for strides of 32 and more)
#endif
#ifndef SPY_KS_REDUCE
#define SPY_KS_REDUCE 0   // 1: a filling buffer is cut by sampled pivots only (no sort) -- measured 1 % slower with 8 drain warps, and
                          // (together with SPY_KS_DEFER) 3-7 % slower with 16 (profiles/r02/probe_stream_y.txt)
#endif
#ifndef SPY_KS_DBUF
#define SPY_KS_DBUF 1     // sparse hand-overs alternate between two TMEM regions: the expansion side runs up to two panels ahead
#endif
#ifndef SPY_KS_DEPTH
#define SPY_KS_DEPTH 1    // batches of KS_U chunks per lane in flight (ring of KS_DEPTH * KS_U chunks per lane); 2 needs SPY_KS_SEARCH
#endif
#ifndef SPY_KS_DEFER
#define SPY_KS_DEFER 0    // 1: raw candidates are evaluated when a selection needs them, not at the end of every panel (two barriers
                          // fewer per panel; measured: no gain -- the expansion side is the critical path, see REDUCE)
#endif
#ifndef SPY_KS_VECQ
#define SPY_KS_VECQ 1    // the drain queues the passing quads of the four tiles of a group at once (0: tile by tile)
#endif
#ifndef SPY_KS_KEEPTAU
#define SPY_KS_KEEPTAU 1  // a validated speculative bound stays the row's bound for the panels that follow
#endif
#ifndef SPY_KS_PREFETCH
#define SPY_KS_PREFETCH 0  // bulk L2 prefetch of the next pass's segments: measured slower (43.1 vs 40.4 ms, profiles/r02)
#endif
constexpr int KS_NT = 1024;
constexpr int KS_D_WARPS = SPY_KS_D_WARPS;     // warps 0..D-1: drain (4 or 8: whole lane quarters)
constexpr int KS_X_WARPS = 32 - KS_D_WARPS;    // warps D..31: expansion side; its first warp also runs the row queue
constexpr int KS_A_WARPS = KS_X_WARPS - 1;     // warps D+1..31 expand
constexpr int KS_X_THREADS = KS_X_WARPS * 32;
constexpr int KS_DT = KS_D_WARPS * 32;
constexpr int KS_U = SPY_KS_U;                 // 16-byte chunks per lane per batch (ring: 16 * KS_U bytes per lane)
#ifndef SPY_KS_SPARSE
#define SPY_KS_SPARSE (SPY_KS_D_WARPS == 16)   // touched-list hand-over of sparse panels (the 16-drain-warp build)
#endif
// Sparse panels (SPY_KS_SPARSE): the expansion lists the slots it touches for the first time; a panel with at most
// KS_LIST_CAP touched slots is handed over as (column, sum) pairs -- KS_SP_J pairs per expansion-side thread, in the TMEM
// lane of that thread -- instead of as a 160 KB snapshot that the drain has to sweep.
#ifndef SPY_KS_CH
#define SPY_KS_CH (SPY_KS_SPARSE ? 512 : 1056)
#endif
// Entries of a target row per pass, in blocks of 32 (the sparse build: the list takes the room).  1056 rather than 1024: a row
// of 1000 +- 32 entries -- the target rows of configs[1..3] -- then fits ONE pass per panel (3.8 % of them do not, against
// 22 %); the few entries behind 1024 used to cost a second, nearly empty pass per panel (8.3k cycles each: +27 % on rows of
// 1040 entries, profiles/r02/probe_ch.txt).  33 blocks is what the shared memory of the 5-panel plan has room for.
constexpr int KS_CH = SPY_KS_CH;
constexpr int KS_NB = KS_CH / 32;                  // staged blocks per pass: lane b keeps block b, lanes 0.. also block 32 + b
constexpr int KS_NBT = KS_NB > 32 ? ((KS_NB + 1) / 2) * 2 : 32;  // block totals per staging buffer
static_assert(KS_CH % 32 == 0 && KS_NB <= KS_NBT && (KS_NB <= 32 || SPY_KS_SEARCH), "pass size");
constexpr int KS_LIST_CAP = SPY_KS_SPARSE ? 6144 : 0;
constexpr int KS_SP_J = (KS_LIST_CAP + KS_X_THREADS - 1) / KS_X_THREADS;  // pairs per thread
constexpr int KS_SP_COLS = 2 * KS_SP_J;                                    // TMEM columns per (lane quarter, sub-group of warps)
static_assert(!SPY_KS_SPARSE || SPY_KS_LOCAL, "the list holds slot offsets");
static_assert(SPY_KS_LOCAL, "the counting form of the panel (UNIT) is written for slot offsets in the chunks");
static_assert(!SPY_KS_SPARSE || (KS_X_WARPS == KS_D_WARPS && KS_A_WARPS <= KS_CH / 32 && 8 * KS_SP_COLS <= 512 && KS_SP_J % 4 == 0),
              "sparse hand-over: drain warp (quarter, sub) reads what expansion-side warp (quarter, sub) wrote");
constexpr int KS_CAP = 1024;                   // candidate buffer (keys); k <= KS_CAP / 2
constexpr int KS_S_WARPS = KS_D_WARPS < 8 ? KS_D_WARPS : 8;  // drain warps that take part in the sample of a first panel
constexpr int KS_QBATCH = KS_D_WARPS > 8 ? 16 : 32;  // queued quads that trigger their per-slot test
constexpr int KS_QCAP = KS_QBATCH + 32;          // quads per drain warp waiting for it (a tile adds up to 32)
constexpr int KS_FLAG_PANEL_END = 1, KS_FLAG_ROW_END = 2, KS_FLAG_STOP = 4;
static_assert(KS_D_WARPS % 4 == 0 && KS_D_WARPS >= 4 && KS_D_WARPS <= 16, "drain warps must cover whole lane quarters");

struct KsPass {   // one pass = up to KS_CH entries of a target row against one panel (shared memory, ring of 4)
    long long aoff;  // first entry of the pass in aexp
    int voff;        // ... and in a_data
    int n;           // entries
    int pn, flags, t, i_out;
};
struct KsQueue {  // cursor of the row queue (shared memory; lane 0 of the queue warp)
    long long tq0;
    int have, i_out, t, a0, len, pn, c0;
};
struct KsMsg {    // what the expansion side hands to the drain with every snapshot (shared memory, double buffered)
    int t, i_out, pn, flags, landed;
    int sparse_n;  // >= 0: the panel was handed over as that many (column, sum) pairs; -1: as a dense snapshot
};

__host__ __device__ constexpr size_t ks_ring_bytes() { return SPY_KS_RING ? (size_t)KS_A_WARPS * 32 * KS_U * 16 * SPY_KS_DEPTH : 0; }
__host__ __device__ constexpr size_t ks_stage_bytes() { return (size_t)2 * (3 * KS_CH * 4 + KS_NBT * 4); }
__host__ __device__ constexpr size_t ks_queue_bytes() { return (size_t)KS_D_WARPS * KS_QCAP * (16 + 4); }
__host__ __device__ constexpr size_t ks_pad_bytes() { return SPY_KS_LOCAL ? 16 : 0; }  // the slot filler pairs are added to
__host__ __device__ constexpr size_t ks_list_bytes() { return (size_t)KS_LIST_CAP * 2; }
__host__ __device__ constexpr size_t ks_fixed_bytes() { return ks_pad_bytes() + ks_ring_bytes() + ks_stage_bytes() + ks_queue_bytes() + (size_t)KS_CAP * 8 + ks_list_bytes(); }

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
// (suspend-time hint: the waiting thread sleeps in the barrier unit until the phase completes or ~20 us pass, instead of
// spinning through the loop below -- the spinning lanes of the idle side took 7 % of all issued instructions, ncu r02v3)
#ifndef SPY_KS_WAIT_HINT
#define SPY_KS_WAIT_HINT 20000
#endif
__device__ __forceinline__ bool ks_mbar_try(unsigned bar, unsigned parity) {
    unsigned ok;
#if SPY_KS_WAIT_HINT
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"((unsigned)SPY_KS_WAIT_HINT) : "memory");
#else
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
    return ok != 0;
}
// bounded wait (about two seconds): a protocol bug must end in an error, not in a hung device
static __device__ __noinline__ void ks_timeout(int *err, int code) {
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
// four tiles (16 TMEM columns) of this warp's lane quarter: r[4 i + c] = slot c of the lane's quad in tile i
__device__ __forceinline__ void ks_tmem_ld16(unsigned taddr, float (&r)[16]) {
    unsigned u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
                   "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void ks_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ks_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- drain side: selection with the threads of the drain warps ---------------------------------------------------
__device__ __forceinline__ void ks_dsync() { ks_bar_sync(4, KS_DT); }

// Bitonic sort (descending) of cand[0, S), S a power of two; the drain warps only.
static __device__ void ks_bitonic_desc(u64 *cand, int S, int dtid) {
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
// One sampling round (see pivot_compact in knn_kernel.cuh): warp 0 sorts 64 strided samples in registers, the pivot sits
// 3 sigma below the sample quantile of the k-th best; the keys above it are compacted IN PLACE (every thread holds its
// keys in registers across the barrier that separates the reads from the writes).  Returns the new n, or -1 when the
// pivot was unlucky (fewer than k keys above it): nothing has been moved then.
static __device__ int ks_pivot_round(u64 *cand, int n, int k, int j, int *s_wsum, u64 *s_pivot, int dtid, u64 *pivot_out = nullptr) {
    constexpr int KPT = (KS_CAP + KS_DT - 1) / KS_DT;  // keys per thread
    const int lane = dtid & 31, w = dtid >> 5;
    if (dtid < 32) {
        u64 k0 = cand[(int)(((long long)dtid * n) >> 6)];
        u64 k1 = cand[(int)(((long long)(dtid + 32) * n) >> 6)];
#pragma unroll
        for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                if (stride == 32) {
                    const u64 hi = k0 > k1 ? k0 : k1, lo = k0 > k1 ? k1 : k0;
                    k0 = hi; k1 = lo;
                } else {
                    const u64 o0 = __shfl_xor_sync(0xffffffffu, k0, stride), o1 = __shfl_xor_sync(0xffffffffu, k1, stride);
                    const bool lower = (dtid & stride) == 0;
                    const bool desc0 = size == 64 || (size == 32 ? true : (dtid & size) == 0);
                    const bool desc1 = size == 64 || (size == 32 ? false : (dtid & size) == 0);
                    k0 = ((k0 > o0) == (lower == desc0)) ? k0 : o0;
                    k1 = ((k1 > o1) == (lower == desc1)) ? k1 : o1;
                }
            }
        }
        const u64 pv = __shfl_sync(0xffffffffu, (j - 1) < 32 ? k0 : k1, (j - 1) & 31);
        if (dtid == 0) *s_pivot = pv;
    }
    ks_dsync();
    const u64 pivot = *s_pivot;
    if (pivot_out) *pivot_out = pivot;
    u64 kk[KPT];
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < KPT; r++) {
        const int i = dtid + r * KS_DT;
        kk[r] = i < n ? cand[i] : 0ull;
        cnt += kk[r] > pivot ? 1 : 0;
    }
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_wsum[w] = inc;
    ks_dsync();  // all keys are in registers; the warp totals are visible
    int base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < KS_D_WARPS; i++) {
        const int v = s_wsum[i];
        if (i < w) base += v;
        total += v;
    }
    if (total >= k) {
        int pos = base + inc - cnt;
#pragma unroll
        for (int r = 0; r < KPT; r++)
            if (kk[r] > pivot) cand[pos++] = kk[r];
    }
    ks_dsync();
    return total >= k ? total : -1;
}
// Bitonic sort (descending) of cand[0, S) with the keys in REGISTERS (thread t holds keys t, t + KS_DT, ...): the steps with
// a stride below 32 are warp shuffles, only the others go through shared memory (6 of the 36 steps of 256 keys) -- a step
// through shared memory costs two barriers of the drain warps, and the drain's barriers are slow because its warps share
// their schedulers with the expansion.
template <int S>
static __device__ __noinline__ void ks_sort_regs(u64 *cand, int dtid) {
    constexpr int KPT = (S + KS_DT - 1) / KS_DT;
    u64 kk[KPT];
#pragma unroll
    for (int r = 0; r < KPT; r++) {
        const int i = r * KS_DT + dtid;
        kk[r] = i < S ? cand[i] : 0ull;
    }
    for (int size = 2; size <= S; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                ks_dsync();  // (the previous readers of cand are done)
#pragma unroll
                for (int r = 0; r < KPT; r++) {
                    const int i = r * KS_DT + dtid;
                    if (i < S) cand[i] = kk[r];
                }
                ks_dsync();
            }
#pragma unroll
            for (int r = 0; r < KPT; r++) {
                const int i = r * KS_DT + dtid;
                if (i < S) {  // (whole warps: S and KS_DT are multiples of 32)
                    const u64 o = stride >= 32 ? cand[i ^ stride] : __shfl_xor_sync(0xffffffffu, kk[r], stride);
                    const bool lower = (i & stride) == 0, desc = ((i & size) == 0) || (size == S);
                    kk[r] = ((kk[r] > o) == (lower == desc)) ? kk[r] : o;
                }
            }
        }
    }
    ks_dsync();
#pragma unroll
    for (int r = 0; r < KPT; r++) {
        const int i = r * KS_DT + dtid;
        if (i < S) cand[i] = kk[r];
    }
    ks_dsync();
}
// Make room in a candidate buffer that is filling up: keep a superset of the best k of the evaluated keys cand[0, n) -- a
// few sampling rounds, each of which drops the keys below a pivot that at least k keys beat -- and raise tau to the last
// pivot, a valid LOWER bound of the k-th best.  No sort; the exact selection runs when the rounds do not free half of the
// buffer, and once at the end of the row.  Returns the new n.
static __device__ int ks_select(u64 *cand, int n, int k, u64 &tau, int *s_live, int *s_wsum, u64 *s_pivot, int dtid);
static __device__ __noinline__ int ks_reduce(u64 *cand, int n, int k, u64 &tau, int *s_live, int *s_wsum, u64 *s_pivot, int dtid) {
    for (int round = 0; round < 4; round++) {
        const float qq = 65.f * (float)k / (float)max(n, 1);
        const int j = (int)ceilf(qq + 3.f * sqrtf(qq) + 1.5f);
        if (n <= 2 * k || n <= 192 || j > 40) break;
        u64 pv = 0ull;
        const int c = ks_pivot_round(cand, n, k, j, s_wsum, s_pivot, dtid, &pv);
        if (c < 0) break;
        n = c;
        if (pv > tau) tau = pv;
    }
    if (n > KS_CAP / 2) n = ks_select(cand, n, k, tau, s_live, s_wsum, s_pivot, dtid);
    return n;
}
// Exact selection among the evaluated keys cand[0, n) (0 = dead): leaves the best m = min(k, live) sorted best-first
// in cand[0, m); returns m and, when m == k, sets tau to the k-th key.  (The reference's heap, s_plus.h:45-59.)
// Sampling rounds shrink the set to a few hundred keys before the bitonic network; results never depend on the samples.
static __device__ int ks_select(u64 *cand, int n, int k, u64 &tau, int *s_live, int *s_wsum, u64 *s_pivot, int dtid) {
    for (int round = 0; round < 3; round++) {
        const float qq = 65.f * (float)k / (float)max(n, 1);
        const int j = (int)ceilf(qq + 3.f * sqrtf(qq) + 1.5f);
        if (n <= 2 * k || n <= 192 || j > 40) break;
        const int c = ks_pivot_round(cand, n, k, j, s_wsum, s_pivot, dtid);
        if (c < 0) break;
        n = c;
    }
#if SPY_KS_REGSORT
    int S = 256;
    while (S < n) S <<= 1;
    for (int i = n + dtid; i < S; i += KS_DT) cand[i] = 0ull;
    if (dtid == 0) *s_live = 0;
    ks_dsync();
    if (S == 256) ks_sort_regs<256>(cand, dtid);
    else if (S == 512) ks_sort_regs<512>(cand, dtid);
    else ks_sort_regs<1024>(cand, dtid);
#else
    int S = 32;
    while (S < n) S <<= 1;
    for (int i = n + dtid; i < S; i += KS_DT) cand[i] = 0ull;
    if (dtid == 0) *s_live = 0;
    ks_dsync();
    ks_bitonic_desc(cand, S, dtid);
#endif
    for (int i = dtid; i < S; i += KS_DT)  // live keys are a prefix: find its end
        if (cand[i] != 0ull && (i == S - 1 || cand[i + 1] == 0ull)) *s_live = i + 1;
    ks_dsync();
    const int m = min(*s_live, k);
    if (m == k) tau = cand[k - 1];
    ks_dsync();
    return m;
}

// -DSPY_KS_TIMING=1: cycle counters per role, summed over the CTAs into 24 u64 at the end of the scratch (development builds)
#ifndef SPY_KS_TIMING
#define SPY_KS_TIMING 0
#endif
#if SPY_KS_TIMING
#define KS_T0(var) long long var = clock64()
#define KS_ACC(id, var) do { const long long _n = clock64(); kt[id] += _n - var; var = _n; } while (0)
#define KS_CNT(id, n) do { kt[id] += (n); } while (0)
#else
#define KS_T0(var) do { } while (0)
#define KS_ACC(id, var) do { } while (0)
#define KS_CNT(id, n) do { } while (0)
#endif

// Block search of the expansion for chunk f of a pass of more than 32 blocks, when f lies behind block 31: the totals of
// blocks 32.. are read from the staging buffer (lane b: block 32 + b).  Returns the block, its word (chunks of the pass
// up to and including it | staged segments << 26) and the word of the block before it.
static __device__ __noinline__ void ks_block_behind_31(const unsigned *cP, unsigned f, unsigned bw, int lane, int &blk, unsigned &wb, unsigned &wp) {
    const unsigned w2 = lane < KS_NB - 32 ? cP[3 * KS_CH + 32 + lane] : 0u;
    unsigned inc2 = w2 & 0x3ffffffu;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {  // (at most 8 blocks behind block 31)
        const unsigned v = __shfl_up_sync(0xffffffffu, inc2, o);
        if (lane >= o) inc2 += v;
    }
    const unsigned p1 = __shfl_sync(0xffffffffu, bw, 31);
    const unsigned bw2 = (inc2 + (p1 & 0x3ffffffu)) | (w2 & 0xfc000000u);
    const int b2 = __ffs(__ballot_sync(0xffffffffu, lane < KS_NB - 32 && (bw2 & 0x3ffffffu) > f)) - 1;
    blk = 32 + b2;
    wb = __shfl_sync(0xffffffffu, bw2, b2);
    const unsigned p2 = __shfl_sync(0xffffffffu, bw2, max(b2 - 1, 0));
    wp = b2 > 0 ? p2 : p1;
}

// The kernel.  KIND selects the drain's pre-filter like in knn_flat_kernel (KIND_RAW / _T / _C / _D / _GEN).
// UNIT: every stored value of A and B is 1.0 (binary=True, s_plus_utils.pyx:301-304): a product is exactly 1, so the panel
// COUNTS with the native integer shared-memory add (ATOMS.ADD instead of the LDS / FADD / ATOMS.CAST.SPIN loop of a float
// add); zero is the "untouched" mark, and the hand-over converts the counts to the floats the drain expects (exact below
// 2^24, i.e. always: a count is bounded by the length of a row).
template <int KIND, bool UNIT = false>
__global__ void __launch_bounds__(KS_NT, 1)
knn_stream_kernel(const __grid_constant__ KnnStreamDev p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const KnnDev &q = p.q;
    float *acc = reinterpret_cast<float *>(smem_raw);
    unsigned char *ptr = smem_raw + (size_t)q.W * sizeof(float) + ks_pad_bytes();
    const unsigned ring32 = (unsigned)__cvta_generic_to_shared(ptr);
    ptr += ks_ring_bytes();
    // staged pass (double buffered): per entry the chunk count prefix INSIDE its block of 32 entries, the first chunk, the
    // value; per block the chunk total
    // (buffer b at stage0 + b * KS_STAGE_WORDS: prefixes [KS_CH], first chunks [KS_CH], values [KS_CH], block totals [KS_NBT];
    // plain pointer arithmetic -- an array of pointers indexed by the buffer number would live in local memory)
    constexpr int KS_STAGE_WORDS = 3 * KS_CH + KS_NBT;
    unsigned *const stage0 = reinterpret_cast<unsigned *>(ptr);
    ptr += 2 * KS_STAGE_WORDS * 4;
    float4 *qx_all = reinterpret_cast<float4 *>(ptr); ptr += (size_t)KS_D_WARPS * KS_QCAP * 16;
    int *qc_all = reinterpret_cast<int *>(ptr); ptr += (size_t)KS_D_WARPS * KS_QCAP * 4;
    u64 *cand = reinterpret_cast<u64 *>(ptr); ptr += (size_t)KS_CAP * 8;
    unsigned short *list = reinterpret_cast<unsigned short *>(ptr);  // SPY_KS_SPARSE: slots touched for the first time in this panel
    (void)list;

    // hand-over n uses barrier pair n & 1 (and message slot n & 1): "full" is posted by the expansion side, "empty" collects
    // the drain warps' releases.  Two pairs, because sparse hand-overs are double buffered in TMEM (SPY_KS_DBUF).
    __shared__ __align__(8) unsigned long long s_full[2], s_empty[2];
    __shared__ __align__(8) u64 s_pivot;
    __shared__ KsPass s_pass[4];
    __shared__ KsQueue s_queue;
    __shared__ KsMsg s_msg[2];
    __shared__ unsigned s_tmem;
    __shared__ int s_cnt, s_overflow, s_live, s_wsum[KS_D_WARPS];
    // entries of `list` (keeps counting past KS_LIST_CAP: the panel is then handed over densely); two counters, used by
    // alternate hand-overs: the one a hand-over consumed is cleared behind that hand-over's last barrier, when every warp
    // has read it, and is appended to again only after the NEXT hand-over
    __shared__ int s_list_n[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#if SPY_KS_TIMING
    long long kt[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    const float sentinel = __uint_as_float(kSentinelBits);
    const float4 sentinel4 = make_float4(sentinel, sentinel, sentinel, sentinel);
    const float blank = UNIT ? 0.f : sentinel;  // an untouched slot of the shared-memory panel
    const float4 blank4 = make_float4(blank, blank, blank, blank);
    // count -> the float the drain expects (UNIT)
    auto as_sum = [&](float raw) __attribute__((always_inline)) -> float {
        const unsigned u = __float_as_uint(raw);
        return u != 0u ? (float)u : sentinel;
    };
    (void)as_sum;
    const unsigned acc32 = (unsigned)__cvta_generic_to_shared(acc);
    const unsigned full32 = (unsigned)__cvta_generic_to_shared(&s_full[0]), empty32 = (unsigned)__cvta_generic_to_shared(&s_empty[0]);  // pair 1: + 8
    const int nT = q.W >> 9;  // tiles of 512 columns (W is a multiple of 2048: whole groups of four tiles)

    for (int i = tid * 4; i < q.W; i += KS_NT * 4) *reinterpret_cast<float4 *>(acc + i) = blank4;
    if (tid == 0) {
        ks_mbar_init(full32, 1); ks_mbar_init(full32 + 8u, 1);
        ks_mbar_init(empty32, KS_D_WARPS); ks_mbar_init(empty32 + 8u, KS_D_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_cnt = 0; s_overflow = 0; s_live = 0; s_list_n[0] = 0; s_list_n[1] = 0;
#if SPY_KS_LOCAL
        // the spare slot filler pairs add to: never "untouched" (the counting form adds 1 per filler: it starts at 1, so that
        // the first filler is not taken for a first touch and listed)
        acc[q.W] = UNIT ? __uint_as_float(1u) : 0.f;
#endif
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
        // expansion side
        // =====================================================================================================
        const int wx = warp - KS_D_WARPS;     // 0..KS_X_WARPS-1; warp 0 of the side runs the row queue instead of expanding
        const bool is_queue = wx == 0;
        const int wa = wx - 1;                // 0..KS_A_WARPS-1 for the expansion warps
        // ---- row queue: lane 0 of the side's first warp; its cursor (the pass AFTER the last one described) lives in shared
        //      memory so that it does not occupy registers of the expansion warps ----
        auto describe_next = [&](KsPass *d) {  // lane 0 of the queue warp only
            KsQueue &c = s_queue;
            if (!c.have) {
                const int slot = atomicAdd(q.work_counter, 1);
                if (slot >= q.n_targets) {
                    d->aoff = 0; d->voff = 0; d->n = 0; d->pn = 0; d->flags = KS_FLAG_STOP; d->t = 0; d->i_out = 0;
                    return;
                }
                c.i_out = q.row_order ? __ldg(q.row_order + slot) : slot;
                c.t = __ldg(q.targets + c.i_out);
                c.a0 = __ldg(q.a_indptr + c.t);
                c.len = __ldg(q.a_indptr + c.t + 1) - c.a0;
                c.tq0 = __ldg(p.toff + c.i_out);
                c.pn = 0; c.c0 = 0; c.have = 1;
                if (c.len <= 0) {  // empty row: one message so that the drain writes its (empty) output
                    d->aoff = 0; d->voff = 0; d->n = 0; d->pn = q.n_panels - 1; d->flags = KS_FLAG_PANEL_END | KS_FLAG_ROW_END;
                    d->t = c.t; d->i_out = c.i_out;
                    c.have = 0;
                    return;
                }
            }
            const int n = min(KS_CH, c.len - c.c0);
            int flags = 0;
            if (c.c0 + n == c.len) {
                flags |= KS_FLAG_PANEL_END;
                if (c.pn == q.n_panels - 1) flags |= KS_FLAG_ROW_END;
            }
            d->aoff = (long long)c.pn * p.E + c.tq0 + c.c0; d->voff = c.a0 + c.c0; d->n = n; d->pn = c.pn; d->flags = flags;
            d->t = c.t; d->i_out = c.i_out;
            c.c0 += n;
            if (c.c0 == c.len) { c.c0 = 0; c.pn++; if (c.pn == q.n_panels) c.have = 0; }
        };
        // ---- staging of a pass: expansion warp wa owns block wa of 32 entries (loads at the start of the current pass, scan
        //      + store at its end); the queue warp, idle otherwise, takes blocks KS_A_WARPS..31 ----
        uint2 g_se = make_uint2(0u, 0u);
        float g_v = 0.f;
        auto stage_block = [&](int buf, int b, uint2 se, float v) {  // scan of the block's chunk counts, store
            const unsigned c = se.y - se.x;
            unsigned inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            unsigned *st = stage0 + buf * KS_STAGE_WORDS;
#if SPY_KS_SEARCH
            // only the segments that hold chunks are staged, packed to the front of the block: every staged segment then
            // starts at a chunk number of its own (the expansion finds a chunk's segment from a bit mask of the starts).
            // Per segment: chunks before it inside the block; (first chunk in `chunks`) - that prefix; the entry's value.
            // Per block: chunk total | staged segments << 26  (a block holds at most 32 x 32768 chunks).
            const unsigned nm = __ballot_sync(0xffffffffu, c != 0u);
            if (c != 0u) {
                const int i = b * 32 + __popc(nm & ((1u << lane) - 1u));
                st[i] = inc - c; st[KS_CH + i] = se.x - (inc - c); st[2 * KS_CH + i] = __float_as_uint(v);
            }
            if (lane == 31) st[3 * KS_CH + b] = inc | ((unsigned)__popc(nm) << 26);
#else
            const int i = b * 32 + lane;
            st[i] = inc - c; st[KS_CH + i] = se.x; st[2 * KS_CH + i] = __float_as_uint(v);
            if (lane == 31) st[3 * KS_CH + b] = inc;
#endif
        };
        auto stage_fetch = [&](const KsPass &d, int b, uint2 &se, float &v) {
            const int i = b * 32 + lane;
            se = make_uint2(0u, 0u); v = 0.f;
            if (i < d.n) {
                se = __ldg(p.aexp + d.aoff + i); v = __ldg(q.a_data + d.voff + i);
#if SPY_KS_PREFETCH
                // the segment's chunks -> L2, a whole pass before the ring asks for them (exact bytes: no line over-fetch)
                if (se.y > se.x) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.chunks + se.x), "r"((se.y - se.x) * 16u) : "memory");
#endif
            }
        };
        auto stage_queue_blocks = [&](const KsPass &d, int buf) {  // the queue warp's share, start to end
            // blocks behind the pass's last entry: their (zero) totals with one store -- a short target row leaves most of the
            // queue warp's blocks empty, and scanning them one by one made it the last warp at the end-of-pass barrier
            const int nb = (d.n + 31) >> 5;
            if (lane >= max(nb, KS_A_WARPS)) stage0[buf * KS_STAGE_WORDS + 3 * KS_CH + lane] = 0u;
            if (KS_NB > 32 && lane < KS_NB - 32 && 32 + lane >= max(nb, KS_A_WARPS)) stage0[buf * KS_STAGE_WORDS + 3 * KS_CH + 32 + lane] = 0u;
            for (int b0 = KS_A_WARPS; b0 < nb; b0 += 3) {
                uint2 se[3];
                float v[3];
#pragma unroll
                for (int r = 0; r < 3; r++) { se[r] = make_uint2(0u, 0u); v[r] = 0.f; if (b0 + r < KS_NB) stage_fetch(d, b0 + r, se[r], v[r]); }
#pragma unroll
                for (int r = 0; r < 3; r++) if (b0 + r < KS_NB) stage_block(buf, b0 + r, se[r], v[r]);
            }
        };

        // ---- expansion warp state ----
        // chunk r of a batch: + r * 512; with SPY_KS_DEPTH == 2 the second batch in flight: + 512 * KS_U
        const unsigned slot32 = ring32 + (unsigned)((max(wa, 0) * 32 * KS_U * SPY_KS_DEPTH + lane) * 16);
        unsigned roff = 0u;  // ring half of the NEXT batch to issue (SPY_KS_DEPTH == 2)
        unsigned f = 0u, fb = 0u, F1 = 0u;  // next chunk to issue, end of the sub-batch, end of the warp's range (uniform)
        unsigned total = 0u;                // chunks of the pass
        float V = 0.f;
        const unsigned *cP = nullptr;  // staged pass
        float vp[KS_U];
        unsigned lp = 0u;
#if SPY_KS_DEPTH == 2
        static_assert(SPY_KS_RING && SPY_KS_SEARCH, "two batches in flight: the ring form with the bit-mask search");
        float vq[KS_U];   // values / live chunks of the YOUNGER batch in flight (vp / lp: the older one)
        unsigned lq = 0u;
        int nfl = 0;      // batches in flight
#pragma unroll
        for (int r = 0; r < KS_U; r++) vq[r] = 0.f;
#endif
#if !SPY_KS_RING
        uint4 nx[KS_U];  // the batch in flight
#pragma unroll
        for (int r = 0; r < KS_U; r++) nx[r] = make_uint4(0u, 0u, 0u, 0u);
#endif
#pragma unroll
        for (int r = 0; r < KS_U; r++) vp[r] = 0.f;
#if SPY_KS_SEARCH
        // The window of an expansion warp is one staged BLOCK (up to 32 segments, all of them non-empty): lane l holds
        // segment l of the block.  A batch covers the chunk numbers [f, f + 32 KS_U); the segments that START inside it set
        // one bit each (distinct, because no staged segment is empty), one REDUX.OR per 32 chunks collects them, and the
        // segment of chunk number f + x is  jprev + popc(starts at or before x),  jprev = the last segment that starts
        // before f.  Two warp reductions per batch instead of ten dependent shuffles of a binary search.
        unsigned bw = 0u;                   // lane b: chunks of the pass up to and including block b | staged segments of block b << 26
        // (blocks 32.. of a pass of more than 32 blocks: their totals are re-read from the staging buffer by the few warps whose
        // chunk range reaches them -- a register for them in every expansion warp costs spills in the loop below)
        unsigned Pe = 0xffffffffu, Dl = 0u; // lane l: first chunk number of segment l of the block; (its first chunk in `chunks`) - Pe
        int jprev = -1;
        auto issue = [&]() {  // one batch of KS_U chunks per lane into the ring
            if (f >= fb) {    // the block that holds chunk f (uniform); f < F1 <= total: it exists
                const unsigned in_first = __ballot_sync(0xffffffffu, (bw & 0x3ffffffu) > f);
                int blk;
                unsigned wb, wp;
                if (KS_NB > 32 && in_first == 0u) {  // behind block 31 (rare: kept out of line, away from the registers of the loop)
                    ks_block_behind_31(cP, f, bw, lane, blk, wb, wp);
                } else {
                    blk = __ffs(in_first) - 1;
                    wb = __shfl_sync(0xffffffffu, bw, blk);
                    wp = __shfl_sync(0xffffffffu, bw, max(blk - 1, 0));
                }
                const unsigned bend = wb & 0x3ffffffu, bo = blk > 0 ? (wp & 0x3ffffffu) : 0u;
                const unsigned *sb = cP + blk * 32 + lane;
                Pe = 0xffffffffu;
                if (lane < (int)(wb >> 26)) { Pe = bo + sb[0]; Dl = sb[KS_CH] - bo; V = __uint_as_float(sb[2 * KS_CH]); }
                fb = min(F1, bend);
                jprev = __popc(__ballot_sync(0xffffffffu, Pe < f)) - 1;
            }
            const unsigned s = Pe - f;  // (segments that start before f, and unused lanes: far beyond the batch)
            const unsigned le = (2u << lane) - 1u;
            int cum = jprev;
            lp = 0u;
#pragma unroll
            for (int r = 0; r < KS_U; r++) {
                unsigned bit;
                asm("shl.b32 %0, %1, %2;" : "=r"(bit) : "r"(1u), "r"(s - 32u * (unsigned)r));  // (shift amounts above 31 give 0)
                const unsigned M = __reduce_or_sync(0xffffffffu, bit);
                const int j = cum + __popc(M & le);
                cum += __popc(M);
                const unsigned fr = f + (unsigned)(r * 32 + lane);
                const unsigned dj = __shfl_sync(0xffffffffu, Dl, j);
                vp[r] = __shfl_sync(0xffffffffu, V, j);
                if (fr < fb) {
#if SPY_KS_RING
                    ks_cp_async16(slot32 + roff + (unsigned)r * 512u, p.chunks + (dj + fr));
#else
                    nx[r] = __ldg(p.chunks + (dj + fr));
#endif
                    lp |= 1u << r;
                }
            }
#if SPY_KS_RING
            ks_cp_commit();
#endif
            jprev = cum;
            f = min(f + 32u * KS_U, fb);
        };
#else
        int j0 = 0, n_cur = 0;
        unsigned boff = 0u;                  // lane b: chunks before block b of the pass
        unsigned Pe = 0u, Pi = 0u, C0 = 0u;  // lane l: segment j0 + l of the pass: first chunk number, end, first chunk in `chunks`
        auto prefix_at = [&](int idx) -> unsigned {  // chunks of the pass before segment idx (all lanes call it together)
            const unsigned o = __shfl_sync(0xffffffffu, boff, min(idx, KS_CH - 1) >> 5);
            return idx >= n_cur ? total : cP[idx] + o;
        };
        auto issue = [&]() {  // one batch of KS_U chunks per lane into the ring
            if (f >= fb) {    // next 32 segments (uniform)
                const int idx = j0 + lane;
                Pe = prefix_at(idx);
                Pi = prefix_at(idx + 1);
                C0 = cP[KS_CH + min(idx, n_cur - 1)];
                V = __uint_as_float(cP[2 * KS_CH + min(idx, n_cur - 1)]);
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
#if SPY_KS_RING
                    ks_cp_async16(slot32 + (unsigned)r * 512u, p.chunks + (c0 + (fr - pe)));
#else
                    nx[r] = __ldg(p.chunks + (c0 + (fr - pe)));
#endif
                    lp |= 1u << r;
                }
            }
#if SPY_KS_RING
            ks_cp_commit();
#endif
            f = min(f + 32u * KS_U, fb);
        };
#endif

        // the panel's shared address, held in a register: as a known constant it is rebuilt from the CTA's rank in front of
        // every add (S2UR / UMOV / UIADD3 / ULEA, a sixth of the instructions of the expansion loop)
        unsigned accl = acc32;
        asm volatile("add.u32 %0, %0, %1;" : "+r"(accl) : "r"((unsigned)q.n_targets >> 31));

        // ---- snapshot bookkeeping (uniform over the warps of the side) ----
        unsigned seq = 0u;  // snapshots / messages posted so far
        bool panel_landed = false, m_landed = false;
        bool prev_dense = false;  // the previous hand-over was a dense snapshot
        auto post = [&](int t, int i_out, int pn, int flags, int landed, int sparse_n) {  // one thread: message + "snapshot full"
            KsMsg *m = &s_msg[seq & 1u];
            m->t = t; m->i_out = i_out; m->pn = pn; m->flags = flags; m->landed = landed; m->sparse_n = sparse_n;
            ks_mbar_arrive(full32 + 8u * (seq & 1u));
        };
#if SPY_KS_SPARSE
        bool listing = true;  // this warp still appends first touches to the list of the open panel (uniform over the warp)
        // a product into the open panel; true: it was the first one to reach its slot (the add returned the "untouched" mark)
        auto add_first = [&](unsigned slot_off, float prod) __attribute__((always_inline)) -> bool {
            if (UNIT) {
                unsigned oldc;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(oldc) : "r"(accl + slot_off) : "memory");
                return oldc == 0u;
            }
            float old;
            asm volatile("atom.shared.add.f32 %0, [%1], %2;" : "=f"(old) : "r"(accl + slot_off), "f"(prod) : "memory");
            return __float_as_uint(old) == kSentinelBits;
        };
#endif
        auto snapshot = [&](const KsPass &dp) {  // dp: the pass that completed the panel
            KS_T0(ts);
            const int m_t = dp.t, m_pn = dp.pn;
#if SPY_KS_SPARSE
            // (all appends of the panel are behind the end-of-pass barrier; nobody appends again before this function's last barrier)
            const int n_list = *reinterpret_cast<volatile int *>(&s_list_n[seq & 1u]);
            const bool sparse = n_list <= KS_LIST_CAP;
#else
            const int n_list = -1;
            const bool sparse = false;
#endif
            // The drain must have released hand-over seq - 2 (its message slot, its barrier pair and -- for a sparse panel --
            // its TMEM region are reused now), and hand-over seq - 1 as well unless both it and this one are sparse: a
            // dense snapshot covers the whole of TMEM.
            const bool dense_now = m_landed && !sparse;
            if (lane == 0) {
                ks_mbar_wait(empty32 + 8u * (seq & 1u), ((seq >> 1) & 1u) ^ 1u, p.err, 2);
                if ((!SPY_KS_DBUF || dense_now || prev_dense) && seq >= 1u)
                    ks_mbar_wait(empty32 + 8u * ((seq & 1u) ^ 1u), ((seq - 1u) >> 1) & 1u, p.err, 2);
            }
            prev_dense = dense_now;
            __syncwarp();
            KS_ACC(4, ts);
            ks_tc_fence_after();
            if (m_landed) {
                const int base = m_pn * q.W;
                if (q.filter_mode == SPY_SEL_MATRIX) {  // erase the row's filtered columns first (s_plus.h:159-172)
                    const int width = min(q.W, q.n_cols - base);
                    const int fs = __ldg(q.f_indptr + m_t), fe = __ldg(q.f_indptr + m_t + 1);
                    for (int x = fs + (tid - KS_DT); x < fe; x += KS_X_THREADS) {
                        const int c = __ldg(q.f_indices + x) - base;
                        if (c >= 0 && c < width) acc[c] = blank;
                    }
                    ks_bar_sync(3, KS_X_THREADS);
                }
                const int sub = wx >> 2;  // the warps of one lane quarter take every (KS_X_WARPS / 4)-th tile
#if SPY_KS_SPARSE
                if (sparse) {  // the listed slots only: (column, sum) pairs into this thread's TMEM lane, the slots reset
                    const int nj = (n_list + KS_X_THREADS - 1) / KS_X_THREADS;
                    for (int j0 = 0; j0 < nj; j0 += 4) {  // four rounds at a time: independent chains, one tcgen05.st
                        unsigned col[4], sl[4];
                        float x[4];
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const int i = (j0 + r) * KS_X_THREADS + (tid - KS_DT);
                            sl[r] = i < n_list ? (unsigned)list[i] : 0xffffffffu;
                        }
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            col[r] = 0xffffffffu; x[r] = sentinel;
                            if (sl[r] != 0xffffffffu) { x[r] = UNIT ? as_sum(acc[sl[r]]) : acc[sl[r]]; col[r] = (unsigned)base + sl[r]; }
                        }
#pragma unroll
                        for (int r = 0; r < 4; r++)
                            if (sl[r] != 0xffffffffu) acc[sl[r]] = blank;
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                                     ::"r"(tmem_q + (unsigned)(sub * KS_SP_COLS + 2 * j0) + (SPY_KS_DBUF ? (seq & 1u) * (unsigned)(4 * KS_SP_COLS) : 0u)), "r"(col[0]), "r"(__float_as_uint(x[0])),
                                     "r"(col[1]), "r"(__float_as_uint(x[1])), "r"(col[2]), "r"(__float_as_uint(x[2])), "r"(col[3]),
                                     "r"(__float_as_uint(x[3])) : "memory");
                    }
                } else
#endif
                for (int T = sub; T < nT; T += KS_X_WARPS / 4) {
                    const unsigned a = acc32 + (unsigned)(512 * T + 128 * (warp & 3) + 4 * lane) * 4u;
                    float4 x = lds128(a);
                    sts128(a, blank4);
                    if (UNIT) x = make_float4(as_sum(x.x), as_sum(x.y), as_sum(x.z), as_sum(x.w));
                    ks_tmem_st4(tmem_q + (unsigned)(4 * T), x);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            ks_tc_fence_before();
            ks_bar_sync(2, KS_X_THREADS);  // every slot of the panel is in TMEM and reset
            if (tid == KS_DT) {
#if SPY_KS_SPARSE
                s_list_n[seq & 1u] = 0;
#endif
                post(dp.t, dp.i_out, dp.pn, dp.flags, m_landed ? 1 : 0, sparse ? n_list : -1);
            }
#if SPY_KS_SPARSE
            listing = true;
#endif
            seq++;
            KS_ACC(3, ts);
        };

        // set-up of a staged pass (buffer buf): the chunk range of this warp and its first batch; true: the warp has chunks
        auto setup = [&](const KsPass &d, int buf) -> bool {
            // chunks before every block of the pass (lane b: block b) and the pass total
            const unsigned *st = stage0 + buf * KS_STAGE_WORDS;
#if SPY_KS_SEARCH
            const unsigned w = st[3 * KS_CH + lane], bt = w & 0x3ffffffu;
#else
            const unsigned bt = st[3 * KS_CH + lane];
#endif
            unsigned inc = bt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            total = __shfl_sync(0xffffffffu, inc, 31);
#if SPY_KS_SEARCH
            bw = inc | (w & 0xfc000000u);
            if (KS_NB > 32) {  // blocks 32..: lane b keeps block 32 + b
                const unsigned w2 = lane < KS_NB - 32 ? st[3 * KS_CH + 32 + lane] : 0u;
                unsigned inc2 = w2 & 0x3ffffffu;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {  // (at most 8 blocks behind block 31)
                    const unsigned v = __shfl_up_sync(0xffffffffu, inc2, o);
                    if (lane >= o) inc2 += v;
                }
                total += __shfl_sync(0xffffffffu, inc2, 7);
            }
            if (is_queue) return false;
            cP = st;
            const unsigned F0 = (unsigned)(((unsigned long long)total * (unsigned)wa) / KS_A_WARPS);
            F1 = (unsigned)(((unsigned long long)total * (unsigned)(wa + 1)) / KS_A_WARPS);
            f = F0; fb = F0;  // (the first issue loads the block of chunk F0)
            if (F0 >= F1) return false;
#else
            boff = inc - bt;
            if (is_queue) return false;
            n_cur = d.n; cP = st;
            const unsigned F0 = (unsigned)(((unsigned long long)total * (unsigned)wa) / KS_A_WARPS);
            F1 = (unsigned)(((unsigned long long)total * (unsigned)(wa + 1)) / KS_A_WARPS);
            f = F0; fb = F0;
            if (F0 >= F1) return false;
            // j0 = the last segment that starts at or before chunk F0: first its block, then inside the block
            const int blk = __popc(__ballot_sync(0xffffffffu, lane * 32 <= n_cur && boff <= F0)) - 1;
            const unsigned v = prefix_at(blk * 32 + lane);
            j0 = blk * 32 + __popc(__ballot_sync(0xffffffffu, blk * 32 + lane <= n_cur && v <= F0)) - 1;
#endif
            roff = 0u;
            issue();
#if SPY_KS_DEPTH == 2
            // a second batch right behind the first: a short pass has all of its chunks in flight at once
            nfl = 1;
            if (f < F1) {
                float va[KS_U];
#pragma unroll
                for (int r = 0; r < KS_U; r++) va[r] = vp[r];
                const unsigned la = lp;
                roff = 512u * KS_U;
                issue();
#pragma unroll
                for (int r = 0; r < KS_U; r++) { vq[r] = vp[r]; vp[r] = va[r]; }
                lq = lp; lp = la;
                roff = 0u;
                nfl = 2;
            }
#endif
            return true;
        };

        // ---- prologue: describe passes 0 and 1, stage pass 0 ----
        if (tid == KS_DT) {
            s_queue.have = 0;
            describe_next(&s_pass[0]);
            if (s_pass[0].flags & KS_FLAG_STOP) { s_pass[1].flags = KS_FLAG_STOP; s_pass[1].n = 0; }
            else describe_next(&s_pass[1]);
        }
        ks_bar_sync(1, KS_X_THREADS);
        {
            const KsPass d0 = s_pass[0];
            if (is_queue) stage_queue_blocks(d0, 0);
            else { stage_fetch(d0, wa, g_se, g_v); stage_block(0, wa, g_se, g_v); }
        }
        ks_bar_sync(1, KS_X_THREADS);
        bool pending = false;
        KS_T0(tp);
        // (A variant that set the next pass up BEFORE the end-of-pass barrier, with the pass descriptors three ahead and an
        // mbarrier per staging buffer, was measured: equal on the probe, 161 vs 153 ms on configs[1] -- not kept.)
        for (int pass = 0;; pass++) {
            const int buf = pass & 1;
            const KsPass d = s_pass[pass & 3];
            const KsPass dn = s_pass[(pass + 1) & 3];
            const bool stop = (d.flags & KS_FLAG_STOP) != 0;
            const bool stage_next = !stop && !(dn.flags & KS_FLAG_STOP);
            bool have = false;
#if !SPY_KS_LOCAL
            const unsigned base = (unsigned)d.pn * (unsigned)q.W;
#endif
            if (!stop) have = setup(d, buf);  // this warp's chunk range of the pass and its first batch
            if (stage_next && !is_queue) stage_fetch(dn, wa, g_se, g_v);  // pass + 1: the loads land during this pass
            if (is_queue && !stop) {
                if (lane == 0) {  // pass + 2
                    if (stage_next) describe_next(&s_pass[(pass + 2) & 3]);
                    else { s_pass[(pass + 2) & 3].flags = KS_FLAG_STOP; s_pass[(pass + 2) & 3].n = 0; }  // (STOP stays STOP)
                }
                __syncwarp();
            }
            KS_ACC(5, tp);
            if (pending) snapshot(s_pass[(pass + 3) & 3]);  // (pass - 1: its slot is rewritten two passes from now)
            if (stop) break;
            KS_ACC(0, tp);  // (the snapshot, also counted in [3] / [4])
            const bool any = have;
            while (have) {
                uint4 pr[KS_U];
                float vc[KS_U];
#if SPY_KS_RING
#if SPY_KS_DEPTH == 2
                if (nfl == 2) asm volatile("cp.async.wait_group 1;" ::: "memory"); else ks_cp_wait_all();
#pragma unroll
                for (int r = 0; r < KS_U; r++) { pr[r] = ks_lds128u(slot32 + roff + (unsigned)r * 512u); vc[r] = vp[r]; }
#else
                ks_cp_wait_all();
#pragma unroll
                for (int r = 0; r < KS_U; r++) { pr[r] = ks_lds128u(slot32 + (unsigned)r * 512u); vc[r] = vp[r]; }
#endif
#else
#pragma unroll
                for (int r = 0; r < KS_U; r++) { pr[r] = nx[r]; vc[r] = vp[r]; }
#endif
                const unsigned lc = lp;
#if SPY_KS_DEPTH == 2
                // the batch just read was the older one: the younger one (if any) takes its place, and the next batch of the
                // warp's range goes into the ring half that has just been read
                const bool two = nfl == 2;
                nfl--;
                if (nfl == 1) {
#pragma unroll
                    for (int r = 0; r < KS_U; r++) vp[r] = vq[r];
                    lp = lq;
                }
                if (f < F1) {
                    float va[KS_U];
#pragma unroll
                    for (int r = 0; r < KS_U; r++) va[r] = vp[r];
                    const unsigned la = lp;
                    issue();  // (into ring half roff; overwrites vp / lp)
                    if (nfl == 1) {
#pragma unroll
                        for (int r = 0; r < KS_U; r++) { vq[r] = vp[r]; vp[r] = va[r]; }
                        lq = lp; lp = la;
                    }
                    nfl++;
                }
                if (two) roff ^= 512u * KS_U;  // (the older batch of the two left in flight sits in the other half)
                const bool more = nfl > 0;
#else
                const bool more = f < F1;
                if (more) issue();
#endif
#if SPY_KS_SPARSE
                bool first[2 * KS_U];
#pragma unroll
                for (int r = 0; r < 2 * KS_U; r++) first[r] = false;
#endif
#pragma unroll
                for (int r = 0; r < KS_U; r++) {
                    if (lc & (1u << r)) {
#if SPY_KS_LOCAL
                        // pr.x / pr.z: byte offset of the slot inside the panel (a filler pair adds 0 to the slot behind it)
#ifdef SPY_KS_NOADDS  // timing experiment: gathers only (results are wrong)
                        if (pr[r].x + pr[r].z == 0x12345u) smem_add_f32(accl, vc[r]);
#else
#if SPY_KS_SPARSE
                        first[2 * r] = add_first(pr[r].x, __fmul_rn(__uint_as_float(pr[r].y), vc[r]));
                        first[2 * r + 1] = add_first(pr[r].z, __fmul_rn(__uint_as_float(pr[r].w), vc[r]));
#else
                        if (UNIT) {
                            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(accl + pr[r].x) : "memory");
                            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(accl + pr[r].z) : "memory");
                        } else {
                            smem_add_f32(accl + pr[r].x, __fmul_rn(__uint_as_float(pr[r].y), vc[r]));
                            smem_add_f32(accl + pr[r].z, __fmul_rn(__uint_as_float(pr[r].w), vc[r]));
                        }
#endif
#endif
#else
                        const unsigned d0 = pr[r].x - base, d1 = pr[r].z - base;
#ifdef SPY_KS_NOADDS  // timing experiment: gathers only (results are wrong)
                        if (d0 + d1 == 0x12345u) smem_add_f32(accl, vc[r]);
#else
                        if (d0 < (unsigned)q.W) smem_add_f32(accl + d0 * 4u, __fmul_rn(__uint_as_float(pr[r].y), vc[r]));
                        if (d1 < (unsigned)q.W) smem_add_f32(accl + d1 * 4u, __fmul_rn(__uint_as_float(pr[r].w), vc[r]));
#endif
#endif
                    }
                }
#if SPY_KS_SPARSE
                // the slots this batch touched for the first time join the panel's list: one reservation per warp and batch
                if (listing) {
                    unsigned fb_[2 * KS_U];
                    int nf = 0;
#pragma unroll
                    for (int r = 0; r < 2 * KS_U; r++) { fb_[r] = __ballot_sync(0xffffffffu, first[r]); nf += __popc(fb_[r]); }
                    if (nf != 0) {
                        int at = 0;
                        if (lane == 0) at = atomicAdd(&s_list_n[seq & 1u], nf);
                        at = __shfl_sync(0xffffffffu, at, 0);
                        if (at + nf > KS_LIST_CAP) listing = false;  // (the count alone tells the hand-over that the list is incomplete)
                        else {
#pragma unroll
                            for (int r = 0; r < 2 * KS_U; r++) {
                                if (first[r]) list[at + __popc(fb_[r] & ((1u << lane) - 1u))] = (unsigned short)(((r & 1) ? pr[r >> 1].z : pr[r >> 1].x) >> 2);
                                at += __popc(fb_[r]);
                            }
                        }
                    }
                }
#endif
                have = more;
            }
            KS_ACC(1, tp);
            if (stage_next) {
                if (is_queue) stage_queue_blocks(dn, buf ^ 1);
                else stage_block(buf ^ 1, wa, g_se, g_v);
            }
            KS_ACC(6, tp);
            const bool landed = ks_bar_or(1, KS_X_THREADS, any);  // all adds of the pass have landed; the next pass is staged
            KS_ACC(2, tp);
            panel_landed |= landed;
            pending = false;
            if (d.flags & KS_FLAG_PANEL_END) {
                if (panel_landed || (d.flags & KS_FLAG_ROW_END)) { pending = true; m_landed = panel_landed; }
                panel_landed = false;
            }
        }
        // no more rows: tell the drain
        if (tid == KS_DT) {
            ks_mbar_wait(empty32 + 8u * (seq & 1u), ((seq >> 1) & 1u) ^ 1u, p.err, 3);
            post(0, 0, 0, KS_FLAG_STOP, 0, -1);
        }
#if SPY_KS_TIMING
        // [0] snapshot  [1] pass body  [2] end-of-pass barrier  [3] snapshot (again)  [4] wait for the drain  [5] pass setup + first issue
        // [6] staging of the next pass; the first expansion warp reports into 0..6, the queue warp into 8..14
        if (lane == 0 && (wx == 1 || wx == 0))
            for (int i = 0; i < 7; i++) atomicAdd(q.phase + (wx == 0 ? 8 : 0) + i, (u64)kt[i]);
#endif
    } else {
        // =====================================================================================================
        // drain side: warp w reads lane quarter w % 4 of the snapshot; with 8 warps the two warps of a quarter alternate
        // over the groups of four tiles
        // =====================================================================================================
        const int dtid = tid;
        constexpr int DPQ = KS_D_WARPS / 4;  // warps per lane quarter
        const int dsub = warp >> 2, quarter = warp & 3;
        float4 *qx = qx_all + warp * KS_QCAP;
        int *qc = qc_all + warp * KS_QCAP;
        const bool filter = !q.exact_only;
        const bool useT = filter && (KIND == KIND_T || (KIND == KIND_GEN && q.l1 != 0.f));
        const bool useC = filter && (KIND == KIND_C || (KIND == KIND_GEN && q.l2 != 0.f));
        const bool useD = filter && (KIND == KIND_D || (KIND == KIND_GEN && q.l3 != 0.f));
        const int nG = nT >> 2;                                // groups of four tiles
        const int nGloc = (nG - dsub + DPQ - 1) / DPQ;         // groups of this warp's sweep
        bool row_open = false;
        SimRow sr = {0.f, 0.f, 0.f};
        FastRow fr = {0.f, 0.f, 0.f, 0.f, 0.f};
        u64 tau = 0ull;
        float lo = reject_bound(q, tau);
        int n_eval = 0;  // cand[0, n_eval) evaluated keys, cand[n_eval, s_cnt) raw candidates
        int n_buf = 0;   // s_cnt as it stood when the previous panel of the row was done (uniform)
#if SPY_KS_HINT
        float hint = 0.f;      // speculative bound (a similarity value) carried from row to row of this CTA; uniform
        bool hint_ok = false;
#endif
        // exact values of the raw candidates (computeSimilarity, s_plus.h:129-156; threshold, s_plus.h:206)
        auto evaluate = [&](int cnt) __attribute__((always_inline)) {
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
        auto select_now = [&]() __attribute__((always_inline)) {  // evaluate what is buffered, keep the best k, raise tau
            ks_dsync();
            const int cnt = min(*reinterpret_cast<volatile int *>(&s_cnt), KS_CAP);
            evaluate(cnt);
            ks_dsync();
#if SPY_KS_REDUCE
            const int m = ks_reduce(cand, cnt, q.k, tau, &s_live, s_wsum, &s_pivot, dtid);
#else
            const int m = ks_select(cand, cnt, q.k, tau, &s_live, s_wsum, &s_pivot, dtid);
#endif
            lo = reject_bound(q, tau);
            n_eval = m;
            if (dtid == 0) { s_cnt = m; s_overflow = 0; }
            ks_dsync();
        };

        KS_T0(td);
        for (unsigned seq = 0u;; seq++) {
            if (lane == 0) ks_mbar_wait(full32 + 8u * (seq & 1u), (seq >> 1) & 1u, p.err, 1);
            __syncwarp();
            KS_ACC(0, td);
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
                n_buf = 0;
            }
#ifdef SPY_KS_NODRAIN  // timing experiment: the drain releases the snapshot unread (results are wrong)
            if (false) {
#else
            if (m.landed) {
#endif
                const int base = m.pn * q.W;
                // One sweep over this warp's share of the snapshot.  lo_s: an extra (speculative) lower bound; sample: only ONE
                // pseudo-randomly chosen tile, every touched slot of it (see the speculative bound below).
                auto sweep = [&](float lo_s, bool sample) __attribute__((always_inline)) {
                int gi = 0, i_res = 0, qn = 0;  // next group of four tiles of this warp's sweep, first tile of it still to do; queued quads
                for (;;) {  // leaves the loop when the warp's share is done; re-entered after an overflow
                    bool overflow = false;
                    const float lo_u = fmaxf(lo, lo_s);  // (lo itself rises when a full buffer forces a selection)
                    const bool flt = filter && !sample;
#if SPY_KS_SPARSE
                    if (m.sparse_n >= 0) {
                        // The panel came as (column, sum) pairs, KS_X_THREADS per round `gi`, one per TMEM lane: every pair is
                        // tested per slot (its Y values are gathered; there is no lane-wise bound for arbitrary columns) and
                        // the survivors are buffered raw with one reservation per warp.  A sample: the first round only.
                        const int nj = (m.sparse_n + KS_X_THREADS - 1) / KS_X_THREADS;
                        const int nj_use = sample ? min(nj, 1) : nj;
                        for (; gi < nj_use; gi += 4) {  // four rounds at a time: one tcgen05.ld, the Y gathers of all four in flight
                            unsigned u[8];
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                         : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                                         : "r"(tmem_q + (unsigned)(dsub * KS_SP_COLS + 2 * gi) + (SPY_KS_DBUF ? (seq & 1u) * (unsigned)(4 * KS_SP_COLS) : 0u)) : "memory");
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            bool keep[4];
                            float yt[4], yc[4], yd[4];
#pragma unroll
                            for (int r = 0; r < 4; r++) {  // (rounds behind the last one hold pairs of earlier panels)
                                keep[r] = gi + r < nj_use && u[2 * r] != 0xffffffffu && u[2 * r + 1] != kSentinelBits;
                                yt[r] = yc[r] = yd[r] = 0.f;
                                if (keep[r] && flt) {
                                    if (useT) yt[r] = __ldg(q.Yt + (int)u[2 * r]);
                                    if (useC) yc[r] = __ldg(q.Yc + (int)u[2 * r]);
                                    if (useD) yd[r] = __ldg(q.Yd + (int)u[2 * r]);
                                }
                            }
                            unsigned bal[4];
                            int tot = 0;
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                if (keep[r] && flt) keep[r] = !slot_rejected<KIND>(q, fr, __uint_as_float(u[2 * r + 1]), lo_u, yt[r], yc[r], yd[r]);
                                bal[r] = __ballot_sync(0xffffffffu, keep[r]);
                                tot += __popc(bal[r]);
                            }
                            if (tot != 0) {
                                int pos = 0;
                                if (lane == 0) pos = atomicAdd(&s_cnt, tot);
                                pos = __shfl_sync(0xffffffffu, pos, 0);
                                if (pos + tot > KS_CAP) {  // does not fit: dead fillers, select, come back to these rounds
                                    for (int i = pos + lane; i < min(pos + tot, KS_CAP); i += 32) cand[i] = 0xffffffffull;
                                    overflow = true;
                                    break;
                                }
#pragma unroll
                                for (int r = 0; r < 4; r++) {
                                    if (keep[r]) cand[pos + __popc(bal[r] & ((1u << lane) - 1u))] = make_raw(__uint_as_float(u[2 * r + 1]), (int)u[2 * r]);
                                    pos += __popc(bal[r]);
                                }
                            }
                        }
                    } else {
#endif
                    // coarse bound of THIS LANE's slots of the panel (it holds quads lane, lane + 128, ... of its quarter): a slot
                    // can only enter the result if  x >= lo * den  and  den >= Dmin + cX * x  with Dmin from the minima of Y over
                    // the lane's columns, i.e.  x * (1 - lo * cX) >= lo * Dmin.  A register per lane: the sweep below needs no
                    // shuffle and no shared memory per tile (the shared-memory pipe is saturated by the expansion's adds).
                    const float g = 1.f - lo_u * fr.cX;
                    const bool coarse = flt && lo_u > 0.f && g > 0.f && fr.cT >= 0.f && fr.cC >= 0.f && fr.cD >= 0.f &&
                                        (!useT || p.ymin_t != nullptr) && (!useC || p.ymin_c != nullptr) && (!useD || p.ymin_d != nullptr);
                    float bound = 0.f;
                    if (coarse) {
                        const int li = m.pn * 128 + quarter * 32 + lane;
                        float dmin = fr.A0;
                        if (useT) dmin = fmaf(fr.cT, __ldg(p.ymin_t + li), dmin);
                        if (useC) dmin = fmaf(fr.cC, __ldg(p.ymin_c + li), dmin);
                        if (useD) dmin = fmaf(fr.cD, __ldg(p.ymin_d + li), dmin);
                        bound = (KIND == KIND_RAW) ? lo_u : (dmin > 0.f ? lo_u * dmin / g : 0.f);
                    }
                    const float lc = lo_u * (KIND == KIND_D ? fr.cD : fr.cC), la = lo_u * fr.A0;
                    // per-slot test of up to 32 queued quads, all lanes busy: one L2 round trip for the batch; false = buffer full
                    auto batch = [&]() __attribute__((always_inline)) -> bool {
                        KS_T0(tb);
                        __syncwarp();
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
                            sm = survivor_mask<KIND>(q, fr, flt, lo_u, lc, la, x, yt, yc, yd);
                        }
                        __syncwarp();  // the queue slots read above are written again by other lanes of the warp (tile())
                        const int c = __popc(sm);
                        int inc = c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int v = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o) inc += v;
                        }
                        const int tot = __shfl_sync(0xffffffffu, inc, 31);
                        if (tot > 0) {
                            int pos = 0;
                            if (lane == 31) pos = atomicAdd(&s_cnt, tot);
                            pos = __shfl_sync(0xffffffffu, pos, 31);
                            if (pos + tot > KS_CAP) {  // does not fit: dead fillers, select, come back (the quads stay queued)
                                for (int i = pos + lane; i < min(pos + tot, KS_CAP); i += 32) cand[i] = 0xffffffffull;
                                KS_ACC(9, tb);
                                return false;
                            }
                            int w = pos + inc - c;
                            if (sm & 1u) cand[w++] = make_raw(x.x, col0);
                            if (sm & 2u) cand[w++] = make_raw(x.y, col0 + 1);
                            if (sm & 4u) cand[w++] = make_raw(x.z, col0 + 2);
                            if (sm & 8u) cand[w++] = make_raw(x.w, col0 + 3);
                        }
                        qn -= nb;
                        KS_CNT(6, 1);
                        KS_ACC(9, tb);
                        return true;
                    };
                    // coarse test of one quad: false = no slot of it can enter the result
                    auto quad_pass = [&](const float4 x) __attribute__((always_inline)) -> bool {
                        // (an untouched slot holds -0.0f: it fails x >= bound for every bound > 0)
                        bool pass = (x.x >= bound) | (x.y >= bound) | (x.z >= bound) | (x.w >= bound);
                        if (bound <= 0.f || !(KIND == KIND_RAW || KIND == KIND_C || KIND == KIND_D)) {
                            // no usable bound for the lane: every touched slot goes on; and where the denominator may be negative
                            // a negative dot product is left to the per-slot test
                            const bool touched = (__float_as_uint(x.x) != kSentinelBits) | (__float_as_uint(x.y) != kSentinelBits) |
                                                 (__float_as_uint(x.z) != kSentinelBits) | (__float_as_uint(x.w) != kSentinelBits);
                            if (bound <= 0.f) pass = touched;
                            else pass = (pass | (x.x < 0.f) | (x.y < 0.f) | (x.z < 0.f) | (x.w < 0.f)) & touched;
                        }
                        return pass;
                    };
                    // one tile: quads that cannot be rejected as a whole join the queue
                    auto tile = [&](int T, const float4 x, const bool pass) __attribute__((always_inline)) {
                        const unsigned bal = __ballot_sync(0xffffffffu, pass);
                        if (pass) {
                            const int e = qn + __popc(bal & ((1u << lane) - 1u));
                            qx[e] = x;
                            qc[e] = base + 512 * T + 128 * quarter + 4 * lane;
                        }
                        qn += __popc(bal);
                        KS_CNT(11, __popc(bal));
                    };
                    while (!overflow && qn >= KS_QBATCH)  // (a round that follows an overflow starts with a full queue)
                        if (!batch()) overflow = true;
                    if (sample) {
                        // (at most 8 warps sample: 8 tiles of 128 slots never overflow the candidate buffer)
                        if (gi == 0 && warp < KS_S_WARPS) {  // tile (h % nGloc, (h >> 16) % 4) of this warp's share: spread over the panel, different per row
                            unsigned h = (unsigned)m.t * 0x9E3779B1u + (unsigned)warp * 0x85EBCA6Bu;
                            h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
                            const int grp = dsub + DPQ * (int)(h % (unsigned)nGloc), i = (int)((h >> 16) & 3u);
                            float xr[16];
                            ks_tmem_ld16(tmem_q + (unsigned)(16 * grp), xr);
                            const float4 x = i == 0 ? make_float4(xr[0], xr[1], xr[2], xr[3]) : i == 1 ? make_float4(xr[4], xr[5], xr[6], xr[7])
                                           : i == 2 ? make_float4(xr[8], xr[9], xr[10], xr[11]) : make_float4(xr[12], xr[13], xr[14], xr[15]);
                            tile(4 * grp + i, x, quad_pass(x));
                        }
                        gi = 1;
                    } else
                    for (; gi < nGloc && !overflow; gi++) {
                        const int grp = dsub + DPQ * gi;
                        float xr[16];
                        KS_T0(tl);
                        ks_tmem_ld16(tmem_q + (unsigned)(16 * grp), xr);
                        KS_ACC(10, tl);
                        // the four tiles of the group are tested at once -- sixteen independent compares and ONE vote when no quad
                        // of the warp passes (the usual case once a bound exists); the drain warps share their schedulers with
                        // six expansion warps each, so that every dependent instruction of the sweep is expensive
                        bool ps[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) ps[i] = quad_pass(make_float4(xr[4 * i], xr[4 * i + 1], xr[4 * i + 2], xr[4 * i + 3]));
#if SPY_KS_VECQ
                        // ... and queued at once when they fit: four votes, the queue positions of all passing quads from the votes
                        // (four independent chains instead of four dependent tile steps)
                        if (i_res == 0) {
                            unsigned bal[4];
#pragma unroll
                            for (int i = 0; i < 4; i++) bal[i] = __ballot_sync(0xffffffffu, ps[i]);
                            const int tot = __popc(bal[0]) + __popc(bal[1]) + __popc(bal[2]) + __popc(bal[3]);
                            if (tot == 0) continue;
                            if (qn + tot <= KS_QCAP) {
                                const unsigned ltm = (1u << lane) - 1u;
                                int off = qn;
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    if (ps[i]) {
                                        const int e = off + __popc(bal[i] & ltm);
                                        qx[e] = make_float4(xr[4 * i], xr[4 * i + 1], xr[4 * i + 2], xr[4 * i + 3]);
                                        qc[e] = base + 512 * (4 * grp + i) + 128 * quarter + 4 * lane;
                                    }
                                    off += __popc(bal[i]);
                                }
                                qn = off;
                                KS_CNT(11, tot);
                                while (qn >= KS_QBATCH && !overflow)
                                    if (!batch()) { overflow = true; i_res = 4; }  // (the group is consumed: resume behind it)
                                if (overflow) break;
                                continue;
                            }
                        }
#else
                        if (i_res == 0 && !__any_sync(0xffffffffu, ps[0] | ps[1] | ps[2] | ps[3])) continue;
#endif
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            if (i >= i_res && !overflow) {
                                tile(4 * grp + i, make_float4(xr[4 * i], xr[4 * i + 1], xr[4 * i + 2], xr[4 * i + 3]), ps[i]);
                                if (qn >= KS_QBATCH && !batch()) { overflow = true; i_res = i + 1; }
                            }
                        }
                        if (overflow) break;
                        i_res = 0;
                    }
                    while (!overflow && qn > 0)
                        if (!batch()) overflow = true;
#if SPY_KS_SPARSE
                    }
#endif
                    if (overflow && lane == 0) s_overflow = 1;
                    ks_dsync();  // every drain warp is done or stopped at a full buffer
                    const bool again = *reinterpret_cast<volatile int *>(&s_overflow) != 0;
                    if (!again) break;
                    KS_ACC(1, td);
                    select_now();
                    KS_ACC(2, td);
                    KS_CNT(5, 1);
                }
                };
                // ---- speculative bound for a row that has none yet (its first non-empty panel) ----
                // Without a bound the sweep floods the candidate buffer and two or three selections are forced before the
                // pre-filter bites.  Instead: SAMPLE one tile per drain warp (every touched slot of it), take the r-th best
                // sample as the bound -- r chosen so that the panel holds k candidates above it with overwhelming probability --
                // and sweep with it.  The snapshot is read-only, so the bound is simply VALIDATED afterwards (k buffered
                // candidates beat it) and the panel swept again without it if it is not: results never depend on the sample.
                bool done = false;
                KS_T0(tsp);
                // sweep with the speculative bound tau_s, evaluate, count the buffered candidates that beat it (uniform)
                auto try_bound = [&](u64 tau_s) __attribute__((always_inline)) -> int {
                    sweep(reject_bound(q, tau_s), false);
                    const int c2 = min(*reinterpret_cast<volatile int *>(&s_cnt), KS_CAP);
                    evaluate(c2);
                    if (dtid == 0) s_live = 0;
                    ks_dsync();
                    int above = 0;
                    for (int i = dtid; i < c2; i += KS_DT) above += cand[i] > tau_s ? 1 : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
                    if (lane == 0 && above) atomicAdd(&s_live, above);
                    ks_dsync();
                    above = *reinterpret_cast<volatile int *>(&s_live);
                    ks_dsync();
                    return above;
                };
                auto forget = [&]() __attribute__((always_inline)) {  // not validated: forget what was buffered; the valid bound is none again
                    if (dtid == 0) s_cnt = 0;
                    n_eval = 0; tau = 0ull; lo = reject_bound(q, tau);
                    ks_dsync();
                };
                // (a sparse panel: worth a guess from 4 rounds of pairs on)
                const bool spec_geom = (SPY_KS_SPARSE && m.sparse_n >= 0) ? m.sparse_n >= 4 * KS_X_THREADS : nG >= 2 * DPQ;
                if (SPY_KS_SPEC && filter && tau == 0ull && n_eval == 0 && n_buf == 0 && spec_geom) {  // (uniform over the drain warps)
#if SPY_KS_HINT
                    // (a) the bound that the first panel of this CTA's previous row validated: consecutive rows of one matrix
                    //     look alike, and a wrong guess only costs the sweep that finds it out
                    if (hint_ok) {
                        const u64 tau_h = (u64)ordered_bits(hint) << 32;
                        const int above = try_bound(tau_h);
                        done = above >= q.k;
                        KS_CNT(12, 1);
                        if (done) {  // keep the bound where about 2 k candidates of a first panel beat it
                            if (above > 3 * q.k) hint += 0.08f * fabsf(hint);
                            else if (2 * above < 3 * q.k) hint -= 0.05f * fabsf(hint);
#if SPY_KS_KEEPTAU
                            // k buffered candidates beat the guess: it is a valid lower bound of the row's k-th best from here
                            // on -- the next panels are swept with it instead of flooding the buffer into a forced selection
                            tau = tau_h; lo = reject_bound(q, tau);
#endif
                        }
                        else { forget(); hint_ok = false; KS_CNT(13, 1); }
                    }
                    if (!done) {
#else
                    {
#endif
                    // (b) a bound from a sample of this panel
                    sweep(0.f, true);
                    const int cnt = min(*reinterpret_cast<volatile int *>(&s_cnt), KS_CAP);
                    evaluate(cnt);
                    ks_dsync();
                    const int width = min(q.W, q.n_cols - base);
                    const float f = (SPY_KS_SPARSE && m.sparse_n >= 0) ? (float)KS_X_THREADS / (float)max(m.sparse_n, 1)
                                                                       : (float)(KS_S_WARPS * 128) / (float)max(width, 1);
                    const float kf = (float)q.k * f;
#ifdef SPY_SPEC_FORCE_RANK  // test builds: a bound that is almost never valid, to exercise the re-sweep
                    const int r_s = SPY_SPEC_FORCE_RANK;
#else
                    const int r_s = (int)ceilf(kf + 3.5f * sqrtf(kf) + 1.5f);
#endif
                    u64 tau_s = 0ull;
                    const bool usable = f < 0.4f && cnt >= 8 * r_s;  // uniform
                    if (usable && ks_select(cand, cnt, r_s, tau_s, &s_live, s_wsum, &s_pivot, dtid) < r_s) tau_s = 0ull;
                    ks_dsync();
                    if (dtid == 0) s_cnt = 0;  // the sample was only read: its slots are all still in the snapshot
                    // (and a bound that a full buffer forced during the sampling is forgotten with the keys it was taken from:
                    // the k-th of them would be swept again and fail the strict test against itself)
                    n_eval = 0; tau = 0ull; lo = reject_bound(q, tau);
                    ks_dsync();
                    if (tau_s != 0ull) {
                        done = try_bound(tau_s) >= q.k;
                        KS_CNT(7, done ? 0 : 1);
                        if (!done) forget();
#if SPY_KS_KEEPTAU
                        if (done) { tau = tau_s; lo = reject_bound(q, tau); }
#endif
#if SPY_KS_HINT
                        if (done) { hint = unordered_bits((unsigned)(tau_s >> 32)); hint_ok = true; }
#ifdef SPY_HINT_STRESS  // test builds: a carried bound that is (almost) never valid, to exercise its failure path
                        hint = 8.f * fabsf(hint) + 1.f;
#endif
#endif
                    }
                    }
                }
                KS_ACC(8, tsp);
                if (!done) sweep(0.f, false);
            }
            KS_ACC(1, td);
#if SPY_KS_DEFER
            // Every path above ends behind a barrier of the drain warps that follows the last append, and nobody appends
            // again before the NEXT snapshot -- which the expansion side can only post after every drain warp's arrival
            // below: the count read here is final and the same in all warps.
            const int cnt = min(*reinterpret_cast<volatile int *>(&s_cnt), KS_CAP);
#endif
            // the snapshot has been read: hand TMEM back (and with it the message slot)
            ks_tc_fence_before();
            __syncwarp();
            if (lane == 0) ks_mbar_arrive(empty32 + 8u * (seq & 1u));
#if SPY_KS_DEFER
            // tighten the bound while the buffer is reasonably full; otherwise the raw candidates wait for the selection
            // that needs them (a panel's end costs no barrier)
            n_buf = cnt;
            if (cnt > KS_CAP / 2 && !(m.flags & KS_FLAG_ROW_END)) { select_now(); n_buf = n_eval; KS_CNT(5, 1); }
            else if (m.flags & KS_FLAG_ROW_END) { evaluate(cnt); ks_dsync(); }
#else
            // evaluate what this panel added; tighten the bound while the buffer is reasonably full
            ks_dsync();
            {
                const int cnt = min(*reinterpret_cast<volatile int *>(&s_cnt), KS_CAP);
                n_buf = cnt;
                if (cnt > KS_CAP / 2 && !(m.flags & KS_FLAG_ROW_END)) { select_now(); n_buf = n_eval; KS_CNT(5, 1); }
                else { evaluate(cnt); ks_dsync(); }
            }
#endif
            KS_ACC(3, td);
            if (m.flags & KS_FLAG_ROW_END) {
                // ---- final selection and slab write (s_plus.h:443-450) ----
                const int n_out = ks_select(cand, n_eval, q.k, tau, &s_live, s_wsum, &s_pivot, dtid);
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
                KS_ACC(4, td);
            }
        }
#if SPY_KS_TIMING
        // drain warp 0 -> 16..22: [0] wait for a snapshot  [1] sweep  [2] selections forced by a full buffer  [3] evaluate / tighten
        // [4] final selection + write  [5] selections  [6] per-slot batches
        // [8] speculative path (sample, bound, sweep, validation)  [9] per-slot batches  [10] tcgen05.ld of the sweeps  [11] quads queued
        // [12] rows that tried the carried bound  [13] ... and failed
        if (tid == 0) for (int i = 0; i < 16; i++) atomicAdd(q.phase + 16 + i, (u64)kt[i]);  // [7] speculative bounds that failed validation
#endif
    }
    ks_tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_tmem) : "memory");
}

typedef void (*knn_stream_kernel_t)(const KnnStreamDev);

}  // inline namespace
}  // namespace spy

// ---- host side (knn_stream.cu, knn_stream_sparse.cu), used by the planner / launcher in knn_kernel.cu -----------
namespace spy {
struct StreamPlan {
    int W, n_panels, cap, drain_warps;
    size_t smem_bytes;
};
// Drain warps of the two builds of the kernel (8 / 16 unless a development build overrides them).
int stream_drain_warps(bool sparse);
// false when the stream engine does not cover the configuration (k too large for its candidate buffer).
// drain_warps selects the build; sp.drain_warps returns it.
bool stream_plan(int k, int n_cols, int panel_width, int max_smem_optin, int drain_warps, StreamPlan &sp);
int64_t stream_scratch_bytes(int n_panels);
int stream_launch(const spy_knn_args &a, const StreamPlan &sp, int kind, int exact_only, int grid, void *scratch,
                  int64_t scratch_bytes, cudaStream_t st);
}  // namespace spy
