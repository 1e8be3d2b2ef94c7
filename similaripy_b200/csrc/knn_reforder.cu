// knn_reforder.cu -- the similarity kernel in the REFERENCE'S OWN ORDER (opt-in: tuning={"tie_mode": "reference"}).
//
// The fast engines resolve exact ties at the k-th value deterministically (larger value first, then smaller column) and add
// the partial products of a column in arbitrary order.  The reference's results depend on order twice: its float sums are
// accumulated entry by entry of the target row (s_plus.h:418-438 / 358-403), and its heap keeps, among candidates that tie
// at the k-th value, those that ARRIVED first in the touched-list order (s_plus.h:45-59, 112-117, 193-215).  This
// kernel restates that order exactly -- slower by design (one barrier per entry of the target row), for callers that need
// the reference's index sets on tie-heavy (binary, count) data:
//   * entries of the target row are expanded ONE AFTER THE OTHER, columns of a B row in parallel (they are distinct), so
//     every column's sum is built in the reference's order: bit-identical float32 values;
//   * a column joins the candidate list when its sum is 0 before the add (s_plus.h:112-117, including the reference's quirk
//     that a sum returning to exactly 0 enlists the column a second time, as a zero-valued candidate);
//   * with column blocking (block_size < n_cols) the row is expanded block by block over B's columns -- which the caller has
//     permuted by popularity exactly like _reorder_columns_by_popularity (s_plus_utils.pyx:493-618) -- so the list is
//     block-major like the reference's stream, and ties compare PERMUTED ids like its heap;
//   * selection: tau = k-th largest value; every candidate above tau is kept; among the ties the heap's behaviour has a
//     closed form: let P be the first k list entries with value >= tau; the ties of P survive except the n smallest ids,
//     n = number of above-tau candidates that arrive after P (each evicts the smallest tied id from a full heap).
#include "knn_kernel.cuh"
#include <algorithm>

namespace spy {

constexpr int RO_NT = 256;

struct ReforderDev {
    KnnDev q;
    int block_size;            // columns per block of the reference's blocked path (>= n_cols: unblocked)
    const int *out_col_map;    // emulation column id -> caller's column id (NULL: identity)
    int list_cap;              // list entries per CTA
    float *acc;                // [grid][n_cols] sums, zero between rows
    unsigned *fpos;            // [grid][n_cols] list position of a column's first enlisting, 0xffffffff between rows
    int *list_col;             // [grid][list_cap]
    unsigned *list_key;        // [grid][list_cap] ordered value bits, 0 = not a candidate
};

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, *total = sum (two barriers)
__device__ __forceinline__ int ro_block_scan(int x, int *s_warp, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    __syncthreads();  // (s_warp may still be read from the previous call)
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < RO_NT / 32; i++) {
        const int v = s_warp[i];
        if (i < w) base += v;
        tot += v;
    }
    *total = tot;
    return base + inc - x;
}

// rank-th (1-based) largest (LARGEST) or smallest 32-bit value among the list entries for which sel(p, value&) holds
template <bool LARGEST, typename SEL>
__device__ unsigned ro_radix_select(int n, int rank, SEL sel, int *s_hist) {
    unsigned prefix = 0u, pmask = 0u;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += RO_NT) s_hist[i] = 0;
        __syncthreads();
        for (int p = threadIdx.x; p < n; p += RO_NT) {
            unsigned v;
            if (sel(p, v) && (v & pmask) == prefix) atomicAdd(&s_hist[(v >> shift) & 255u], 1);
        }
        __syncthreads();
        // every thread walks the 256 bins (uniform result)
        int bin = 0, acc = 0;
        if (LARGEST) { for (bin = 255; bin > 0; bin--) { if (acc + s_hist[bin] >= rank) break; acc += s_hist[bin]; } }
        else { for (bin = 0; bin < 255; bin++) { if (acc + s_hist[bin] >= rank) break; acc += s_hist[bin]; } }
        rank -= acc;
        prefix |= (unsigned)bin << shift;
        pmask |= 255u << shift;
        __syncthreads();
    }
    return prefix;
}

__global__ void __launch_bounds__(RO_NT) knn_reforder_kernel(const __grid_constant__ ReforderDev p) {
    const KnnDev &q = p.q;
    extern __shared__ __align__(16) unsigned char ro_smem[];
    u64 *s_out = reinterpret_cast<u64 *>(ro_smem);  // [pow2(k)] final sort
    __shared__ int s_warp[RO_NT / 32], s_hist[256], s_row, s_n;
    const int tid = threadIdx.x;
    float *acc = p.acc + (size_t)blockIdx.x * q.n_cols;
    unsigned *fpos = p.fpos + (size_t)blockIdx.x * q.n_cols;
    int *lcol = p.list_col + (size_t)blockIdx.x * p.list_cap;
    unsigned *lkey = p.list_key + (size_t)blockIdx.x * p.list_cap;
    const int bs = max(1, min(p.block_size, max(q.n_cols, 1)));
    const int n_blocks = (q.n_cols + bs - 1) / bs;
    int kp = 32;
    while (kp < q.k) kp <<= 1;

    for (;;) {
        if (tid == 0) s_row = atomicAdd(q.work_counter, 1);
        __syncthreads();
        const int i_out = s_row;
        __syncthreads();
        if (i_out >= q.n_targets) break;
        const int t = __ldg(q.targets + i_out);
        const int a0 = __ldg(q.a_indptr + t), a1 = __ldg(q.a_indptr + t + 1);
        // ---- expansion in the reference's order: block by block, entry by entry (s_plus.h:350-438) ----
        int n_list = 0;
        for (int b = 0; b < n_blocks; b++) {
            const int c_lo = b * bs, c_hi = min(q.n_cols, c_lo + bs);
            for (int j = a0; j < a1; j++) {
                const int u = __ldg(q.a_indices + j);
                const float v = __ldg(q.a_data + j);
                int s = __ldg(q.b_indptr + u), e = __ldg(q.b_indptr + u + 1);
                if (n_blocks > 1) {  // the part of the (sorted) row inside the block (std::lower_bound, s_plus.h:381-394)
                    s = lower_bound_dev(q.b_indices, s, e, c_lo);
                    e = lower_bound_dev(q.b_indices, s, e, c_hi);
                }
                for (int x0 = s; x0 < e; x0 += RO_NT) {
                    const int x = x0 + tid;
                    int c = -1, flag = 0;
                    if (x < e) {
                        c = __ldg(q.b_indices + x);
                        const float old = acc[c];
                        flag = old == 0.f ? 1 : 0;  // SparseMatrixMultiplier::add (s_plus.h:112-117)
                        acc[c] = __fadd_rn(old, __fmul_rn(v, __ldg(q.b_data + x)));
                    }
                    int total;
                    const int pos = n_list + ro_block_scan(flag, s_warp, &total);
                    if (flag && pos < p.list_cap) {
                        lcol[pos] = c;
                        if (fpos[c] == 0xffffffffu) fpos[c] = (unsigned)pos;
                    }
                    n_list = min(n_list + total, p.list_cap);
                }
            }
        }
        __syncthreads();
        // ---- values of the candidates (foreach, s_plus.h:193-215): filter / target selectors, computeSimilarity, threshold ----
        SimRow sr;
        sr.Xt = (q.l1 != 0.f) ? __ldg(q.Xt + t) : 0.f;
        sr.Xc = (q.l2 != 0.f) ? __ldg(q.Xc + t) : 0.f;
        sr.Xd = (q.l3 != 0.f) ? __ldg(q.Xd + t) : 0.f;
        int fs = 0, fe = 0, ts = 0, te = 0;
        if (q.filter_mode == SPY_SEL_MATRIX) { fs = __ldg(q.f_indptr + t); fe = __ldg(q.f_indptr + t + 1); }
        if (q.target_mode == SPY_SEL_MATRIX) { ts = __ldg(q.t_indptr + t); te = __ldg(q.t_indptr + t + 1); }
        for (int x = tid; x < n_list; x += RO_NT) {
            const int c = lcol[x];
            const float xy = fpos[c] == (unsigned)x ? acc[c] : 0.f;  // a second enlisting of a column finds its sum already drained
            bool ok = true;
            if (q.filter_mode == SPY_SEL_MATRIX) { const int w = lower_bound_dev(q.f_indices, fs, fe, c); ok = !(w < fe && __ldg(q.f_indices + w) == c); }
            if (ok && q.target_mode == SPY_SEL_MATRIX) { const int w = lower_bound_dev(q.t_indices, ts, te, c); ok = w < te && __ldg(q.t_indices + w) == c; }
            unsigned key = 0u;
            if (ok) {
                const float val = similarity_value(q, sr, xy, q.l1 != 0.f ? __ldg(q.Yt + c) : 0.f, q.l2 != 0.f ? __ldg(q.Yc + c) : 0.f,
                                                   q.l3 != 0.f ? __ldg(q.Yd + c) : 0.f);
                if (val >= q.thr) key = ordered_bits(val);
            }
            lkey[x] = key;
        }
        __syncthreads();
        for (int x = tid; x < n_list; x += RO_NT) { const int c = lcol[x]; acc[c] = 0.f; fpos[c] = 0xffffffffu; }  // clean for the next row
        // ---- how many candidates, tau ----
        int mine = 0;
        for (int x = tid; x < n_list; x += RO_NT) mine += lkey[x] != 0u ? 1 : 0;
        int n_cand;
        ro_block_scan(mine, s_warp, &n_cand);
        unsigned tau = 0u;   // keep everything with key >= tau ...
        int p_star = n_list; // ... among list positions <= p_star ...
        unsigned id_th = 0u; // ... ties only with id > id_th (when drop > 0)
        int drop = 0;
        if (n_cand > q.k) {
            tau = ro_radix_select<true>(n_list, q.k, [&](int x, unsigned &v) { v = lkey[x]; return v != 0u; }, s_hist);
            int g_mine = 0, z_mine = 0;
            for (int x = tid; x < n_list; x += RO_NT) { const unsigned v = lkey[x]; g_mine += v > tau ? 1 : 0; z_mine += v == tau ? 1 : 0; }
            int g, z;
            ro_block_scan(g_mine, s_warp, &g);
            ro_block_scan(z_mine, s_warp, &z);
            if (g + z > q.k) {  // more ties than places: the heap's arrival rule decides (see the header)
                // p_star = list position of the k-th entry with key >= tau
                int seen = 0;
                p_star = -1;
                for (int x0 = 0; x0 < n_list && p_star < 0; x0 += RO_NT) {
                    const int x = x0 + tid;
                    const int f = (x < n_list && lkey[x] >= tau) ? 1 : 0;
                    int total;
                    const int excl = ro_block_scan(f, s_warp, &total);
                    if (tid == 0) s_n = -1;
                    __syncthreads();
                    if (f && seen + excl + 1 == q.k) s_n = x;
                    __syncthreads();
                    if (s_n >= 0) p_star = s_n;
                    seen += total;
                }
                int gp_mine = 0;
                for (int x = tid; x <= p_star; x += RO_NT) gp_mine += lkey[x] > tau ? 1 : 0;
                int gp;
                ro_block_scan(gp_mine, s_warp, &gp);
                drop = g - gp;  // above-tau candidates that arrive when the heap is full: each evicts the smallest tied id
                if (drop > 0)
                    id_th = ro_radix_select<false>(p_star + 1, drop, [&](int x, unsigned &v) { v = (unsigned)lcol[x]; return lkey[x] == tau; }, s_hist);
            }
        }
        // ---- gather the kept candidates, order them best-first, write the slab row (s_plus.h:443-450) ----
        for (int x = tid; x < kp; x += RO_NT) s_out[x] = 0ull;
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (int x = tid; x < n_list; x += RO_NT) {
            const unsigned v = lkey[x];
            if (v == 0u || v < tau) continue;
            bool keep = v > tau || tau == 0u || n_cand <= q.k;
            if (!keep) keep = x <= p_star && (drop == 0 || (unsigned)lcol[x] > id_th);  // a tie
            if (keep) {
                const int c = lcol[x];
                const int oc = p.out_col_map ? __ldg(p.out_col_map + c) : c;
                const int w = atomicAdd(&s_n, 1);
                if (w < kp) s_out[w] = ((u64)v << 32) | (u64)(0xffffffffu - (unsigned)oc);
            }
        }
        __syncthreads();
        const int n_out = min(s_n, q.k);
        for (int size = 2; size <= kp; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = tid; i < (kp >> 1); i += RO_NT) {
                    const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1)), hi = lo + stride;
                    const bool desc = ((lo & size) == 0) || (size == kp);
                    const u64 a = s_out[lo], b = s_out[hi];
                    if ((a < b) == desc) { s_out[lo] = b; s_out[hi] = a; }
                }
                __syncthreads();
            }
        const size_t o = (size_t)i_out * (size_t)q.k;
        for (int j = tid; j < q.k; j += RO_NT) {
            int col = 0, row = 0;
            float val = 0.f;
            if (j < n_out) {
                const u64 key = s_out[j];
                col = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
                val = unordered_bits((unsigned)(key >> 32));
                row = t;
            }
            q.out_cols[o + j] = col;
            q.out_vals[o + j] = val;
            if (q.out_rows) q.out_rows[o + j] = row;
        }
        if (tid == 0 && q.out_counts) q.out_counts[i_out] = n_out;
        __syncthreads();
    }
}

__global__ void ro_fill_u32(unsigned *p, size_t n, unsigned v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

static int ro_grid(int n_targets, int n_cols) {
    // bounded by the scratch it implies: 16 bytes per column per CTA
    long long g = std::min<long long>(n_targets, (long long)kB200SmCount * 4);
    const long long by_mem = std::max<long long>(1, (1ll << 31) / std::max<long long>(16ll * std::max(n_cols, 1), 1));
    return (int)std::max<long long>(1, std::min(g, by_mem));
}

}  // namespace spy

using namespace spy;

extern "C" {

int64_t spy_knn_reforder_scratch_bytes(const spy_knn_args *args) {
    if (!args) return SPY_ERR_INVALID;
    const int grid = ro_grid(std::max(args->n_targets, 1), args->n_cols);
    const int64_t nc = std::max(args->n_cols, 1), cap = 2 * nc + 1024;
    return 256 + (int64_t)grid * (nc * 8 + cap * 8);
}

int spy_knn_topk_reforder_dev(const spy_knn_args *args, int32_t block_size, const int32_t *out_col_map, void *scratch,
                              int64_t scratch_bytes, void *stream) {
    SPY_REQUIRE(args != nullptr, "args is NULL");
    const spy_knn_args &a = *args;
    SPY_REQUIRE(a.k >= 1 && a.k <= 4096, "tie_mode='reference' supports k <= 4096 (got %d)", a.k);
    if (a.n_targets <= 0) return SPY_OK;
    SPY_REQUIRE(a.out_cols && a.out_values, "output slab pointers are NULL");
    SPY_REQUIRE(scratch && scratch_bytes >= spy_knn_reforder_scratch_bytes(args), "scratch too small");
    SPY_REQUIRE(a.l1 == 0.f || (a.Xtversky && a.Ytversky), "l1 != 0 needs Xtversky/Ytversky");
    SPY_REQUIRE(a.l2 == 0.f || (a.Xcosine && a.Ycosine), "l2 != 0 needs Xcosine/Ycosine");
    SPY_REQUIRE(a.l3 == 0.f || (a.Xdepop && a.Ydepop), "l3 != 0 needs Xdepop/Ydepop");
    const int grid = ro_grid(a.n_targets, a.n_cols);
    const size_t nc = (size_t)std::max(a.n_cols, 1), cap = 2 * nc + 1024;
    ReforderDev p;
    KnnDev &d = p.q;
    memset(&p, 0, sizeof(p));
    d.n_targets = a.n_targets; d.targets = a.targets;
    d.a_indptr = a.a_indptr; d.a_indices = a.a_indices; d.a_data = a.a_data;
    d.b_indptr = a.b_indptr; d.b_indices = a.b_indices; d.b_data = a.b_data;
    d.n_cols = a.n_cols;
    d.Xt = a.Xtversky; d.Yt = a.Ytversky; d.Xc = a.Xcosine; d.Yc = a.Ycosine; d.Xd = a.Xdepop; d.Yd = a.Ydepop;
    d.a1 = a.a1; d.l1 = a.l1; d.l2 = a.l2; d.l3 = a.l3; d.t1 = a.t1; d.t2 = a.t2;
    d.stab = a.stabilized_shrink; d.bayes = a.bayesian_shrink; d.thr = a.threshold;
    d.k = a.k;
    d.filter_mode = a.filter_mode; d.f_indptr = a.filter_indptr; d.f_indices = a.filter_indices;
    d.target_mode = a.target_mode; d.t_indptr = a.target_indptr; d.t_indices = a.target_indices;
    d.out_rows = a.out_rows; d.out_cols = a.out_cols; d.out_vals = a.out_values; d.out_counts = a.out_counts;
    unsigned char *sc = reinterpret_cast<unsigned char *>(scratch);
    d.work_counter = reinterpret_cast<int *>(sc);
    p.block_size = block_size > 0 ? block_size : std::max(a.n_cols, 1);
    p.out_col_map = out_col_map;
    p.list_cap = (int)std::min<size_t>(cap, 0x7fffffff);
    p.acc = reinterpret_cast<float *>(sc + 256);
    p.fpos = reinterpret_cast<unsigned *>(sc + 256 + (size_t)grid * nc * 4);
    p.list_col = reinterpret_cast<int *>(sc + 256 + (size_t)grid * nc * 8);
    p.list_key = reinterpret_cast<unsigned *>(sc + 256 + (size_t)grid * nc * 8 + (size_t)grid * cap * 4);
    cudaStream_t st = as_stream(stream);
    SPY_CUDA_OK(cudaMemsetAsync(sc, 0, 256 + (size_t)grid * nc * 4, st));
    ro_fill_u32<<<kB200SmCount * 4, 256, 0, st>>>(p.fpos, (size_t)grid * nc, 0xffffffffu);
    SPY_LAUNCH_OK();
    int kp = 32;
    while (kp < a.k) kp <<= 1;
    knn_reforder_kernel<<<grid, RO_NT, (size_t)kp * 8, st>>>(p);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

}  // extern "C"
