// knn_stream.cu -- host side of the stream engine (knn_stream_kernel.cuh): the per-call tables it reads
// (B as padded 16-byte chunks, the chunk range of every (entry of a target row, panel)), planning and launch.
#include "knn_stream_kernel.cuh"
#include <algorithm>

namespace spy {

// Stream layout of B: every (row u, panel p) SEGMENT -- the entries of B[u,:] whose columns fall into panel p -- is stored as
// whole 16-byte chunks of two (column, value) pairs, at chunks [seg[u * P + p], seg[u * P + p + 1]).
//   counts[u * P + p] = ceil(len(u, p) / 2)
__global__ void chunk_counts_kernel(int b_rows, const int *__restrict__ b_indptr, const int *__restrict__ split, int split_stride,
                                    int n_panels, int *__restrict__ counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)b_rows * n_panels) return;
    const int u = (int)(i / n_panels), pn = (int)(i % n_panels);
    int s, e;
    if (n_panels == 1) { s = b_indptr[u]; e = b_indptr[u + 1]; }
    else { s = split[(size_t)u * split_stride + pn]; e = split[(size_t)u * split_stride + pn + 1]; }
    counts[i] = (e - s + 1) >> 1;
}

// One warp per row of B.  Inside a segment the pairs are ordered by SHARED-MEMORY BANK of their accumulator slot (column
// mod 32): the expansion adds pair 0 of 32 consecutive chunks with one instruction and pair 1 with the next, i.e. the
// even and the odd positions of 64 consecutive pairs -- in bank order each of the two sets touches a bank at most
// ceil(m / 2) times when the segment holds m columns of that bank, instead of the ~3.5-way conflicts of 32 random banks
// (the kernel is bound by shared-memory wavefronts: profiles/r02/knn_stream_v1_ncu_summary.txt).  Order inside a segment is
// free: a segment is a set of (column, value) to be added.  Groups of 64 pairs are ordered independently; an odd segment
// ends with the filler pair (0xffffffff, 0), which no panel accepts.
__global__ void pad_chunks_kernel(int b_rows, const int *__restrict__ b_indptr, const int *__restrict__ b_indices,
                                  const float *__restrict__ b_data, const int *__restrict__ split, int split_stride,
                                  int n_panels, const int *__restrict__ seg, uint2 *__restrict__ pairs_out, int W) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long u = warp0; u < b_rows; u += n_warps) {
        for (int pn = 0; pn < n_panels; pn++) {
            int s, e;
            if (n_panels == 1) { s = b_indptr[u]; e = b_indptr[u + 1]; }
            else { s = split[(size_t)u * split_stride + pn]; e = split[(size_t)u * split_stride + pn + 1]; }
            uint2 *out = pairs_out + 2 * (size_t)seg[u * n_panels + pn];
#if SPY_KS_LOCAL
            // (byte offset of the accumulator slot inside the panel, value); the filler pair adds 0 to the spare slot behind
            // the panel: the kernel's expansion loop has no test per product at all
            const unsigned shift = 2u, base = (unsigned)pn * (unsigned)W;
            const uint2 filler = make_uint2((unsigned)W * 4u, 0u);
#else
            const unsigned shift = 0u, base = 0u;
            const uint2 filler = make_uint2(0xffffffffu, 0u);  // a column no panel accepts
#endif
            for (int g0 = s; g0 < e; g0 += 64) {  // groups of 64 pairs: two per lane
                const int q0 = g0 + lane, q1 = g0 + 32 + lane;
                const bool v0 = q0 < e, v1 = q1 < e;
                uint2 p0 = filler, p1 = filler;
                if (v0) p0 = make_uint2(((unsigned)b_indices[q0] - base) << shift, __float_as_uint(b_data[q0]));
                if (v1) p1 = make_uint2(((unsigned)b_indices[q1] - base) << shift, __float_as_uint(b_data[q1]));
                const int k0 = v0 ? (int)((p0.x >> shift) & 31u) : 32, k1 = v1 ? (int)((p1.x >> shift) & 31u) : 32;
                int pos0 = 0, pos1 = 0;  // counting sort by bank: pairs of smaller banks first, then by position
                for (int b = 0; b < 32; b++) {
                    const unsigned m0 = __ballot_sync(0xffffffffu, k0 == b), m1 = __ballot_sync(0xffffffffu, k1 == b);
                    const int c = __popc(m0) + __popc(m1);
                    if (b < k0) pos0 += c;
                    if (b < k1) pos1 += c;
                    if (b == k0) pos0 += __popc(m0 & lt);
                    if (b == k1) pos1 += __popc(m0) + __popc(m1 & lt);
                }
                if (v0) out[(g0 - s) + pos0] = p0;
                if (v1) out[(g0 - s) + pos1] = p1;
            }
            if (((e - s) & 1) && lane == 0) out[e - s] = filler;
        }
    }
}

__global__ void row_lengths_kernel(int n_targets, const int *__restrict__ targets, const int *__restrict__ a_indptr,
                                   int *__restrict__ len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_targets) {
        const int t = targets[i];
        len[i] = a_indptr[t + 1] - a_indptr[t];
    }
}

// one warp per target row: lane j of a step owns one entry (u = its column) and writes the chunk range of B[u,:]
// inside every panel -- for a fixed panel the 32 lanes write 256 consecutive bytes
__global__ void build_aexp_kernel(int n_targets, const int *__restrict__ targets, const int *__restrict__ a_indptr,
                                  const int *__restrict__ a_indices, const int *__restrict__ seg, int n_panels,
                                  const long long *__restrict__ toff, long long E, uint2 *__restrict__ aexp) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp0; i < n_targets; i += n_warps) {
        const int t = targets[i];
        const int a0 = a_indptr[t], len = a_indptr[t + 1] - a0;
        const long long tq0 = toff[i];
        for (int j = lane; j < len; j += 32) {
            const int *sp = seg + (size_t)a_indices[a0 + j] * n_panels;
            int s = sp[0];
            for (int p = 0; p < n_panels; p++) {
                const int e = sp[p + 1];
                aexp[(long long)p * E + tq0 + j] = make_uint2((unsigned)s, (unsigned)e);
                s = e;
            }
        }
    }
}

}  // namespace spy

#include "knn_stream_impl.inc"

// the two builds of the kernel (this translation unit: `a`; knn_stream_sparse.cu: `b`)
namespace spy {
bool stream_plan_b(int k, int n_cols, int panel_width, int max_smem_optin, StreamPlan &sp);
int stream_drain_warps_b();
int stream_launch_b(const spy_knn_args &a, const StreamPlan &sp, int kind, int exact_only, int grid, void *scratch,
                    int64_t scratch_bytes, cudaStream_t st);

int stream_drain_warps(bool sparse) { return sparse ? stream_drain_warps_b() : stream_drain_warps_a(); }
bool stream_plan(int k, int n_cols, int panel_width, int max_smem_optin, int drain_warps, StreamPlan &sp) {
    if (drain_warps == stream_drain_warps_b() && drain_warps != stream_drain_warps_a())
        return stream_plan_b(k, n_cols, panel_width, max_smem_optin, sp);
    return stream_plan_a(k, n_cols, panel_width, max_smem_optin, sp);
}
int64_t stream_scratch_bytes(int n_panels) { return stream_scratch_bytes_a(n_panels); }  // (the same for both builds)
int stream_launch(const spy_knn_args &a, const StreamPlan &sp, int kind, int exact_only, int grid, void *scratch,
                  int64_t scratch_bytes, cudaStream_t st) {
    if (sp.drain_warps == stream_drain_warps_b() && sp.drain_warps != stream_drain_warps_a())
        return stream_launch_b(a, sp, kind, exact_only, grid, scratch, scratch_bytes, st);
    return stream_launch_a(a, sp, kind, exact_only, grid, scratch, scratch_bytes, st);
}
}  // namespace spy

using namespace spy;

extern "C" {

int spy_knn_chunk_counts_dev(int32_t b_rows, const int32_t *b_indptr, const int32_t *b_split, int32_t split_stride,
                             int32_t n_panels, int32_t *counts, void *stream) {
    if (b_rows <= 0) return SPY_OK;
    SPY_REQUIRE(b_indptr && counts && n_panels >= 1, "chunk_counts: bad arguments");
    SPY_REQUIRE(n_panels == 1 || (b_split && split_stride >= n_panels + 1), "chunk_counts: n_panels > 1 needs b_split");
    const long long total = (long long)b_rows * n_panels;
    chunk_counts_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(b_rows, b_indptr, b_split, split_stride,
                                                                                       n_panels, counts);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_knn_pad_chunks_dev(int32_t b_rows, const int32_t *b_indptr, const int32_t *b_indices, const float *b_data,
                           const int32_t *b_split, int32_t split_stride, int32_t n_panels, const int32_t *chunk_indptr,
                           void *chunks_out, void *stream, int32_t panel_width) {
    if (b_rows <= 0) return SPY_OK;
    SPY_REQUIRE(b_indptr && chunk_indptr && chunks_out && n_panels >= 1, "pad_chunks: bad arguments");
    SPY_REQUIRE(panel_width > 0 && panel_width % 2048 == 0, "pad_chunks: panel_width of the stream plan missing");
    SPY_REQUIRE(n_panels == 1 || (b_split && split_stride >= n_panels + 1), "pad_chunks: n_panels > 1 needs b_split");
    const long long warps = std::min<long long>(b_rows, (long long)kB200SmCount * 64);
    pad_chunks_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
        b_rows, b_indptr, b_indices, b_data, b_split, split_stride, n_panels, chunk_indptr, reinterpret_cast<uint2 *>(chunks_out),
        panel_width);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_knn_row_lengths_dev(int32_t n_targets, const int32_t *targets, const int32_t *a_indptr, int32_t *len, void *stream) {
    if (n_targets <= 0) return SPY_OK;
    SPY_REQUIRE(targets && a_indptr && len, "row_lengths: NULL pointer");
    row_lengths_kernel<<<(n_targets + 255) / 256, 256, 0, as_stream(stream)>>>(n_targets, targets, a_indptr, len);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_knn_build_aexp_dev(const spy_knn_args *args, void *stream) {
    SPY_REQUIRE(args != nullptr, "args is NULL");
    const spy_knn_args &a = *args;
    if (a.n_targets <= 0 || a.n_entries <= 0) return SPY_OK;
    SPY_REQUIRE(a.n_panels >= 1, "build_aexp: launch plan missing (spy_knn_plan)");
    SPY_REQUIRE(a.targets && a.a_indptr && a.a_indices && a.b_chunk_indptr && a.toff && a.aexp, "build_aexp: NULL pointer");
    const long long warps = std::min<long long>(a.n_targets, (long long)kB200SmCount * 64);
    build_aexp_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
        a.n_targets, a.targets, a.a_indptr, a.a_indices, a.b_chunk_indptr, a.n_panels,
        reinterpret_cast<const long long *>(a.toff), a.n_entries, reinterpret_cast<uint2 *>(const_cast<void *>(a.aexp)));
    SPY_LAUNCH_OK();
    return SPY_OK;
}

}  // extern "C"
