// knn_kernel.cu -- the hot path: CSR x CSR row-expansion SpGEMM with the similarity
// denominator fused in and per-row top-k selection, for sm_100a (B200).
//
// Replaces s_plus::compute_similarities_parallel<int,float> (reference
// similaripy/cython_code/s_plus.h:265-453) behind the C ABI in include/similaripy_b200.h.
//
// This file is the host side: the planner (engine choice, panel width, launch shape), the C-ABI entry points and the
// launcher of the two kernel generations.  Design:
//   * one persistent CTA per SM; CTAs pull target rows from an atomic queue (the reference's
//     `omp for schedule(dynamic)`, s_plus.h:337);
//   * the output columns are cut into panels of `panel_width` columns; a panel of fp32 partial sums lives in shared
//     memory (the reference's `sums` buffer, s_plus.h:91, at shared-memory instead of L2-cache scale, s_plus.h:305-311);
//   * panel boundaries inside every sorted row of B are precomputed once (b_split) -- the reference does a
//     std::lower_bound per (target row, block, B row), s_plus.h:381-394;
//   * accumulation is a shared-memory float add (LDS / FADD / ATOMS.CAST.SPIN, profiles/microbench/accum_bench_r01.txt);
//     "touched" is encoded in the accumulator itself: slots start at -0.0f, and (-0.0f) + x == x, so a slot whose
//     bits are still 0x80000000 was never written (the reference keeps a touched-list, s_plus.h:112-117);
//   * STREAM engine (knn_stream_kernel.cuh, default): expansion and drain run on different warps of the CTA; the chunks
//     of a pass are numbered through and cut into equal ranges per warp, streamed through a cp.async ring; a complete
//     panel is handed to the drain through tensor memory (dense snapshot, or (column, sum) pairs for sparse panels);
//   * FLAT engine (knn_kernel.cuh, round 1): a group of G lanes owns one entry of the target row at a time and streams the
//     run of its B row inside the panel; expansion and drain alternate inside the CTA.  It keeps what the stream engine
//     does not cover (matrix-mode target_cols, exact-only similarities, k > 512);
//   * the drain applies filter / target selectors, computeSimilarity (s_plus.h:129-156) with the same operation order
//     and no FMA contraction, the threshold test (s_plus.h:206), and feeds a running-threshold candidate buffer that is
//     cut to the best k whenever it fills and once at the end of the row;
//   * exact ties at the k-th value: larger value first, then smaller column id (identical sums only -- the float sums
//     themselves depend on the order in which the adds land).
#include "knn_stream_kernel.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace spy {

// (column, value) pairs of B packed into one 8-byte word per stored entry: one LDG.64 per product
// instead of two LDG.32 from two arrays, and half the sector over-fetch on short segments.
__global__ void pack_pairs_kernel(long long nnz, const int *__restrict__ indices, const float *__restrict__ data,
                                  uint2 *__restrict__ pairs) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride)
        pairs[q] = make_uint2((unsigned)indices[q], __float_as_uint(data[q]));
}

// out[b] = min of y over columns [128 b, 128 b + 128): the drain's coarse bound (one warp per block)
__global__ void block_min_kernel(int n, const float *__restrict__ y, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int blk = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (blk * 128 >= n) return;
    float m = __int_as_float(0x7f800000);
    for (int i = blk * 128 + lane; i < min(n, blk * 128 + 128); i += 32) m = fminf(m, y[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) out[blk] = m;
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
    int threads, ctas_per_sm, W, n_panels, split_stride, cap, group, engine;
    bool cand_smem;
    size_t smem_bytes;
    StreamPlan stream;
};

static int similarity_kind(const spy_knn_args &a, int &exact_only);

// Which kernel generation runs the call: the stream engine (knn_stream_kernel) covers every configuration except
// matrix-mode target columns, the exact-only similarities (a1 != 1 or a bayesian shrink: no division-free pre-filter)
// and k beyond what its candidate buffer holds; those, and explicit requests, take the flat engine.
// SPY_ENGINE=flat|stream overrides the automatic choice (kernel experiments).
static int choose_engine(const spy_knn_args &a, int max_smem_optin, StreamPlan &sp) {
    int exact_only = 0;
    similarity_kind(a, exact_only);
    // Which build of the stream kernel: 8 drain warps + 24 on the expansion side, or 16 + 16 for target rows with few scalar
    // products per panel, where the sweep of the panel and the selections are the critical path (configs[4]-shaped
    // operands, 4e3 products per panel: 34.2 vs 40.0 ms; configs[3]-shaped, 1.9e4: 33.1 vs 28.9 ms; profiles/r02).
    // args.group = 8 / 16 asks for one of them (spy_knn_plan returns the choice there); SPY_KS_DRAIN=8|16 overrides.
    int drain = (a.engine == SPY_ENGINE_STREAM && (a.group == stream_drain_warps(false) || a.group == stream_drain_warps(true))) ? a.group : 0;
    if (drain == 0) {
        static const int env = [] { const char *e = getenv("SPY_KS_DRAIN"); return e ? atoi(e) : 0; }();
        if (env == stream_drain_warps(false) || env == stream_drain_warps(true)) drain = env;
    }
    if (drain == 0) {
        drain = stream_drain_warps(false);
        if (a.a_nnz > 0 && a.b_nnz > 0 && a.a_rows > 0 && a.b_rows > 0) {
            const int panels_est = std::max(1, (std::max(a.n_cols, 1) + 40959) / 40960);
            const double per_panel = ((double)a.a_nnz / a.a_rows) * ((double)a.b_nnz / a.b_rows) / panels_est;
            if (per_panel < 5120.0) drain = stream_drain_warps(true);  // (its touched lists hold 6144 slots per panel)
        }
    }
    const bool eligible = !exact_only && a.target_mode != SPY_SEL_MATRIX && (a.threads == 0 || a.threads == 1024) &&
                          stream_plan(a.k, a.n_cols, a.engine == SPY_ENGINE_STREAM ? a.panel_width : 0, max_smem_optin, drain, sp);
    int want = a.engine;
    if (want == SPY_ENGINE_AUTO) {
        static const int env = [] {
            const char *e = getenv("SPY_ENGINE");
            if (e && !strcmp(e, "flat")) return SPY_ENGINE_FLAT;
            if (e && !strcmp(e, "stream")) return SPY_ENGINE_STREAM;
            return SPY_ENGINE_AUTO;
        }();
        want = env != SPY_ENGINE_AUTO ? env : SPY_ENGINE_DEFAULT;
        // (Round 2 kept short rows on the flat engine; with the 16-drain-warp build the stream engine is ahead there too:
        // 31.1 vs 40.9 ms at 1e3 products per panel, 34.2 vs 44.1 ms at 4e3.)
        if (want == SPY_ENGINE_STREAM && !eligible) want = SPY_ENGINE_FLAT;
    }
    if (want == SPY_ENGINE_STREAM && !eligible) return -1;
    return want;
}

static int make_plan(const spy_knn_args &a, int device, Plan &pl) {
    DeviceInfo di = device_info(device);
    if (a.engine != SPY_ENGINE_AUTO && a.engine != SPY_ENGINE_FLAT && a.engine != SPY_ENGINE_STREAM) {
        set_error("engine must be 0 (auto), 1 (flat) or 2 (stream), got %d", a.engine);
        return SPY_ERR_INVALID;
    }
    pl.engine = choose_engine(a, di.max_smem_optin, pl.stream);
    if (pl.engine < 0) {
        set_error("the stream engine does not cover this configuration (matrix-mode target_cols, a1 != 1, bayesian shrink, "
                  "threads != 1024, k > 512 or a panel_width that is not a multiple of 2048)");
        return SPY_ERR_UNSUPPORTED;
    }
    if (pl.engine == SPY_ENGINE_STREAM) {
        pl.threads = KS_NT; pl.ctas_per_sm = 1; pl.cap = pl.stream.cap; pl.cand_smem = true; pl.group = pl.stream.drain_warps;
        pl.W = pl.stream.W; pl.n_panels = pl.stream.n_panels; pl.smem_bytes = pl.stream.smem_bytes;
        int stride = pl.n_panels + 1;
        if (stride <= 8) { int q = 1; while (q < stride) q <<= 1; stride = q; }
        pl.split_stride = (pl.n_panels > 1) ? stride : 0;
        return SPY_OK;
    }
    pl.threads = a.threads ? a.threads : 1024;
#ifdef SPY_WITH_768
    const bool threads_ok = pl.threads == 512 || pl.threads == 768 || pl.threads == 1024;
#else
    const bool threads_ok = pl.threads == 512 || pl.threads == 1024;  // the 768-thread kernels are an experiment build
#endif
    if (!threads_ok) {
        set_error("threads must be 512 or 1024 (got %d)", pl.threads);
        return SPY_ERR_INVALID;
    }
    pl.ctas_per_sm = pl.threads == 512 ? 2 : 1;  // 512 x 2 and 1024 x 1: 64 registers per thread; 768 x 1: 80
    pl.cap = std::max(2048, next_pow2(2 * std::max(a.k, 1)));
    pl.cand_smem = (size_t)pl.cap * 8 <= 65536;
    // staged target-row chunk: 8 bytes per thread (doubles as the selection's scratch)
    const size_t fixed = (size_t)stage_entries(pl.threads) * 8 + (pl.cand_smem ? (size_t)pl.cap * 8 : 0);
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
    // lanes per B-row segment: a batch is 2 * G * unroll pairs (16-byte gathers of 2 pairs, `unroll` per lane before
    // the first add); measured best when a batch covers about two thirds of the mean segment (entries of a B row
    // inside one panel), so that a typical segment takes two nearly full batches
    int G = a.group;
    if (G == 0) {
        G = 8;
        if (a.b_nnz > 0 && a.b_rows > 0) {
            const double seg = (double)a.b_nnz / ((double)a.b_rows * pl.n_panels);
            const double want = seg * 0.65 / (2 * unroll_for(pl.threads));
            G = want <= 2.9 ? 4 : want <= 11.4 ? 8 : want <= 22.7 ? 16 : 32;  // 4-lane groups only pay below ~18-entry segments
        }
    }
    if (G != 4 && G != 8 && G != 16 && G != 32) {
        set_error("group must be 4, 8, 16 or 32 (got %d)", G);
        return SPY_ERR_INVALID;
    }
    pl.group = G;
    return SPY_OK;
}

static knn_kernel_t pick_kernel(int threads, int kind, bool cand_smem, int group) {
    switch (group) {
    case 4: return pick_kernel_g4(threads, kind, cand_smem);
    case 16: return pick_kernel_g16(threads, kind, cand_smem);
    case 32: return pick_kernel_g32(threads, kind, cand_smem);
    default: return pick_kernel_g8(threads, kind, cand_smem);
    }
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

static int64_t block_min_bytes(int n_cols) { return (((int64_t)(std::max(n_cols, 1) + 127) / 128) * 4 + 255) / 256 * 256; }

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
    args->group = pl.group;
    args->engine = pl.engine;
    return SPY_OK;
}

int64_t spy_knn_scratch_bytes(const spy_knn_args *args, int device) {
    if (!args) return SPY_ERR_INVALID;
    Plan pl;
    if (make_plan(*args, device, pl) != SPY_OK) return SPY_ERR_INVALID;
    if (pl.engine == SPY_ENGINE_STREAM) return stream_scratch_bytes(pl.n_panels);
    int64_t bytes = 256 + block_min_bytes(args->n_cols);  // work counter (+ phase counters), per-block minima of Y
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
    SPY_REQUIRE(a.engine == SPY_ENGINE_FLAT || a.engine == SPY_ENGINE_STREAM, "launch plan missing: call spy_knn_plan first");
    SPY_REQUIRE(a.engine == SPY_ENGINE_STREAM || a.b_pairs != nullptr, "b_pairs is NULL: pack B with spy_knn_pack_pairs_dev first");
    SPY_REQUIRE(a.engine == SPY_ENGINE_STREAM || a.n_panels == 1 || a.b_split != nullptr, "n_panels > 1 needs b_split (spy_knn_build_split_dev)");
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
    if (pl.engine == SPY_ENGINE_STREAM) {
        SPY_REQUIRE(pl.W == a.panel_width && pl.n_panels == a.n_panels, "plan fields were changed after spy_knn_plan");
        int exact_only = 0;
        const int kind = similarity_kind(a, exact_only);
        return stream_launch(a, pl.stream, kind, exact_only, grid, scratch, scratch_bytes, as_stream(stream));
    }
    const int64_t bm_bytes = block_min_bytes(a.n_cols);
    int64_t need = 256 + bm_bytes + (pl.cand_smem ? 0 : (int64_t)grid * pl.cap * 8);
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
    d.k = a.k; d.cap = pl.cap; d.group = pl.group;
    d.filter_mode = a.filter_mode; d.f_indptr = a.filter_indptr; d.f_indices = a.filter_indices;
    d.target_mode = a.target_mode; d.t_indptr = a.target_indptr; d.t_indices = a.target_indices;
    d.out_rows = a.out_rows; d.out_cols = a.out_cols; d.out_vals = a.out_values; d.out_counts = a.out_counts;
    d.work_counter = reinterpret_cast<int *>(scratch);
    d.phase = reinterpret_cast<u64 *>(reinterpret_cast<unsigned char *>(scratch) + 128);
    d.cand_global = reinterpret_cast<u64 *>(reinterpret_cast<unsigned char *>(scratch) + 256 + bm_bytes);

    cudaStream_t st = as_stream(stream);
    SPY_CUDA_OK(cudaMemsetAsync(scratch, 0, 256, st));
    // per-block minima of the one Y vector the drain's coarse bound uses (cosine family: Yc, depop only: Yd)
    d.y_block_min = nullptr;
    const float *ysrc = (kind == KIND_C) ? a.Ycosine : (kind == KIND_D) ? a.Ydepop : nullptr;
    if (ysrc != nullptr && a.n_cols > 0) {
        float *bm = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(scratch) + 256);
        const int blocks = (a.n_cols + 127) / 128;
        block_min_kernel<<<(blocks * 32 + 255) / 256, 256, 0, st>>>(a.n_cols, ysrc, bm);
        SPY_LAUNCH_OK();
        d.y_block_min = bm;
    }
    knn_kernel_t kern = pick_kernel(pl.threads, kind, pl.cand_smem, pl.group);
    SPY_CUDA_OK(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    // Experiment switch (off unless SPY_L2_PERSIST=<MB> is set; to be measured): keep the split table -- one random
    // 32-byte sector per (entry, panel), re-read by every target row that contains the entry -- resident in L2
    // against the 1.6 GB stream of pairs, through a per-launch access-policy window.
    static const long persist_mb = [] { const char *e = getenv("SPY_L2_PERSIST"); return e ? atol(e) : 0L; }();
    const void *table = (pl.n_panels > 1) ? (const void *)a.b_split : (const void *)a.b_indptr;
    if (persist_mb > 0 && table != nullptr) {
        cudaDeviceProp prop;
        SPY_CUDA_OK(cudaGetDeviceProperties(&prop, device));
        const size_t carve = std::min((size_t)persist_mb << 20, (size_t)prop.persistingL2CacheMaxSize);
        const size_t bytes = (pl.n_panels > 1) ? (size_t)a.b_rows * a.split_stride * 4 : ((size_t)a.b_rows + 1) * 4;
        const size_t window = std::min(bytes, (size_t)prop.accessPolicyMaxWindowSize);
        if (carve > 0 && window > 0) {
            SPY_CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeAccessPolicyWindow;
            attr.val.accessPolicyWindow.base_ptr = const_cast<void *>(table);
            attr.val.accessPolicyWindow.num_bytes = window;
            attr.val.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)window);
            attr.val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(pl.threads); cfg.dynamicSmemBytes = pl.smem_bytes; cfg.stream = st;
            cfg.attrs = &attr; cfg.numAttrs = 1;
            SPY_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, d));
            count_launch();
            return SPY_OK;
        }
    }
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
    a.group = host_args->group;
    a.b_nnz = b_nnz;
    a.a_nnz = a_nnz;
    a.engine = host_args->engine;
    rc = spy_knn_plan(&a, device);
    if (rc != SPY_OK) { cleanup(); return rc; }
    if (a.n_panels > 1) {
        // the panel split points need ascending columns inside every row of B; the reference accepts unsorted rows on
        // its unblocked path (s_plus.h:418-438), so sort the device copy rather than require it of the caller
        rc = spy_csr_sort_rows_dev(a.b_rows, a.b_indptr, (int32_t *)a.b_indices, (float *)a.b_data, nullptr);
        if (rc != SPY_OK) { cleanup(); return rc; }
        if (!dev_alloc(&d_split, (size_t)a.b_rows * a.split_stride * 4)) { cleanup(); return SPY_ERR_NOMEM; }
        rc = spy_knn_build_split_dev(a.b_rows, a.b_indptr, a.b_indices, a.panel_width, a.n_panels, a.split_stride,
                                     (int32_t *)d_split, nullptr);
        if (rc != SPY_OK) { cleanup(); return rc; }
        a.b_split = (const int32_t *)d_split;
    }
    if (a.engine == SPY_ENGINE_STREAM) {
        // the stream engine's tables: B as padded 16-byte chunks, the chunk range of every (entry, panel)
        void *d_cnt = nullptr, *d_cptr = nullptr, *d_tmp = nullptr, *d_chunks = nullptr, *d_len = nullptr, *d_toff = nullptr, *d_aexp = nullptr;
        const int64_t n_scan = std::max<int64_t>(std::max(a.b_rows, a.n_targets), 1);
        const int64_t n_seg = (int64_t)a.b_rows * a.n_panels;
        if (!dev_alloc(&d_cnt, (size_t)std::max(n_scan, n_seg) * 4) || !dev_alloc(&d_cptr, ((size_t)n_seg + 1) * 4) ||
            !dev_alloc(&d_tmp, (size_t)spy_scan_tmp_bytes(std::max(n_scan, n_seg))) || !dev_alloc(&d_toff, ((size_t)a.n_targets + 1) * 8)) { cleanup(); return SPY_ERR_NOMEM; }
        d_len = d_cnt;
        rc = spy_knn_chunk_counts_dev(a.b_rows, a.b_indptr, a.b_split, a.split_stride, a.n_panels, (int32_t *)d_cnt, nullptr);
        if (rc == SPY_OK) rc = spy_exclusive_scan_i32_dev(n_seg, (const int32_t *)d_cnt, (int32_t *)d_cptr, d_tmp, nullptr);
        int32_t n_chunks = 0;
        if (rc == SPY_OK && cudaMemcpy(&n_chunks, (int32_t *)d_cptr + n_seg, 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = SPY_ERR_CUDA;
        if (rc != SPY_OK) { cleanup(); return rc; }
        if (!dev_alloc(&d_chunks, ((size_t)std::max(n_chunks, 1)) * 16)) { cleanup(); return SPY_ERR_NOMEM; }
        rc = spy_knn_pad_chunks_dev(a.b_rows, a.b_indptr, a.b_indices, a.b_data, a.b_split, a.split_stride, a.n_panels,
                                    (const int32_t *)d_cptr, d_chunks, nullptr, a.panel_width);
        if (rc == SPY_OK) rc = spy_knn_row_lengths_dev(a.n_targets, a.targets, a.a_indptr, (int32_t *)d_len, nullptr);
        if (rc == SPY_OK) rc = spy_exclusive_scan_i64_dev(a.n_targets, (const int32_t *)d_len, (int64_t *)d_toff, d_tmp, nullptr);
        int64_t n_entries = 0;
        if (rc == SPY_OK && cudaMemcpy(&n_entries, (int64_t *)d_toff + a.n_targets, 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = SPY_ERR_CUDA;
        if (rc != SPY_OK) { cleanup(); return rc; }
        if (!dev_alloc(&d_aexp, (size_t)std::max<int64_t>(n_entries, 1) * a.n_panels * 8)) { cleanup(); return SPY_ERR_NOMEM; }
        a.b_chunk_indptr = (const int32_t *)d_cptr; a.b_chunks = d_chunks; a.toff = (const int64_t *)d_toff;
        a.n_entries = n_entries; a.aexp = d_aexp;
        rc = spy_knn_build_aexp_dev(&a, nullptr);
        if (rc != SPY_OK) { cleanup(); return rc; }
    } else {
        void *d_pairs = nullptr;
        if (!dev_alloc(&d_pairs, ((size_t)std::max(b_nnz, 1) + 1) * 8)) { cleanup(); return SPY_ERR_NOMEM; }
        rc = spy_knn_pack_pairs_dev(b_nnz, a.b_indices, a.b_data, d_pairs, nullptr);
        if (rc != SPY_OK) { cleanup(); return rc; }
        a.b_pairs = d_pairs;
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
