// host_api.cu -- host-pointer entry points of the C ABI beyond spy_knn_topk_host: the multi-GPU variant
// and the in-place CSR normalizers exactly as the reference's Cython functions receive their arguments
// (normalization.pyx:97-102, 200-208, 260-271: shape, data, indices, indptr of a scipy CSR matrix).
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace spy;

namespace {

// Pageable host memory -> device through two pinned staging buffers: the copy of chunk i + 1 into pinned memory (several
// host threads) overlaps the DMA of chunk i.  cudaMemcpyAsync from pageable memory stages too, but single-threaded.
struct Staging {
    static constexpr size_t kChunk = 16u << 20;
    void *pinned[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    int device = -1;
    bool ensure(int dev) {
        if (pinned[0] && device == dev) return true;
        release();
        for (int b = 0; b < 2; b++) {
            if (cudaHostAlloc(&pinned[b], kChunk, cudaHostAllocDefault) != cudaSuccess) { release(); return false; }
            if (cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming) != cudaSuccess) { release(); return false; }
        }
        device = dev;
        return true;
    }
    void release() {
        for (int b = 0; b < 2; b++) {
            if (pinned[b]) cudaFreeHost(pinned[b]);
            if (done[b]) cudaEventDestroy(done[b]);
            pinned[b] = nullptr; done[b] = nullptr;
        }
        device = -1;
        cudaGetLastError();
    }
    ~Staging() { release(); }
};
thread_local Staging g_staging;

void parallel_memcpy(void *dst, const void *src, size_t bytes, int n_threads) {
    if (n_threads <= 1 || bytes < (4u << 20)) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> workers;
    const size_t part = (bytes / n_threads + 4095) & ~(size_t)4095;
    for (int t = 1; t < n_threads; t++) {
        const size_t off = part * t;
        if (off >= bytes) break;
        workers.emplace_back([=]() { memcpy((char *)dst + off, (const char *)src + off, std::min(part, bytes - off)); });
    }
    memcpy(dst, src, std::min(part, bytes));
    for (auto &w : workers) w.join();
}

struct DevBuf {  // device copy of a host array, freed on scope exit
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int up(const void *host, size_t bytes) {
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) { set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); return SPY_ERR_NOMEM; }
        if (host) {
            e = cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { set_error("upload failed: %s", cudaGetErrorString(e)); return SPY_ERR_CUDA; }
        }
        return SPY_OK;
    }
};

size_t val_size(int val_dtype) { return val_dtype == SPY_F64 ? 8 : 4; }
size_t idx_size(int idx_dtype) { return idx_dtype == SPY_I64 ? 8 : 4; }

int64_t read_index(const void *arr, int idx_dtype, int64_t i) {
    return idx_dtype == SPY_I64 ? ((const int64_t *)arr)[i] : (int64_t)((const int32_t *)arr)[i];
}

// shared body of the three normalizer entries: upload, run `body` on device pointers, download the values
template <typename F>
int with_device_csr(int64_t n_rows, void *data, int val_dtype, const void *indices, const void *indptr, int idx_dtype,
                    int device, bool need_indices, F body) {
    SPY_REQUIRE(n_rows >= 0, "n_rows must be >= 0");
    SPY_REQUIRE(val_dtype == SPY_F32 || val_dtype == SPY_F64, "val_dtype must be SPY_F32 or SPY_F64");
    SPY_REQUIRE(idx_dtype == SPY_I32 || idx_dtype == SPY_I64, "idx_dtype must be SPY_I32 or SPY_I64");
    if (n_rows == 0) return SPY_OK;
    SPY_REQUIRE(data && indptr && (!need_indices || indices), "NULL pointer");
    const int64_t nnz = read_index(indptr, idx_dtype, n_rows);
    SPY_CUDA_OK(cudaSetDevice(device));
    DevBuf d_data, d_idx, d_ptr;
    int rc = d_data.up(data, (size_t)nnz * val_size(val_dtype));
    if (rc == SPY_OK) rc = d_ptr.up(indptr, (size_t)(n_rows + 1) * idx_size(idx_dtype));
    if (rc == SPY_OK && need_indices) rc = d_idx.up(indices, (size_t)nnz * idx_size(idx_dtype));
    if (rc != SPY_OK) return rc;
    rc = body(d_data.p, d_idx.p, d_ptr.p);
    if (rc != SPY_OK) return rc;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && nnz > 0) e = cudaMemcpy(data, d_data.p, (size_t)nnz * val_size(val_dtype), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("normalizer run failed: %s", cudaGetErrorString(e)); return SPY_ERR_CUDA; }
    return SPY_OK;
}

}  // namespace

extern "C" {

int spy_h2d_staged(void *dst_dev, const void *src_host, int64_t bytes, int device, void *stream) {
    if (bytes <= 0) return SPY_OK;
    SPY_REQUIRE(dst_dev && src_host, "h2d_staged: NULL pointer");
    SPY_CUDA_OK(cudaSetDevice(device));
    cudaStream_t st = as_stream(stream);
    if (!g_staging.ensure(device)) {  // no pinned memory to be had: let the driver stage
        SPY_CUDA_OK(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, st));
        return SPY_OK;
    }
    const int n_threads = (int)std::max(1u, std::min(4u, std::thread::hardware_concurrency() / 2));
    int b = 0;
    for (size_t off = 0; off < (size_t)bytes; off += Staging::kChunk, b ^= 1) {
        const size_t n = std::min(Staging::kChunk, (size_t)bytes - off);
        SPY_CUDA_OK(cudaEventSynchronize(g_staging.done[b]));  // the buffer's previous DMA has finished (no-op the first time)
        parallel_memcpy(g_staging.pinned[b], (const char *)src_host + off, n, n_threads);
        SPY_CUDA_OK(cudaMemcpyAsync((char *)dst_dev + off, g_staging.pinned[b], n, cudaMemcpyHostToDevice, st));
        SPY_CUDA_OK(cudaEventRecord(g_staging.done[b], st));
    }
    return SPY_OK;
}

int spy_normalize_rows_host(int norm, int64_t n_rows, void *data, int val_dtype, const void *indptr, int idx_dtype, int device) {
    SPY_REQUIRE(norm >= 0 && norm <= 2, "norm must be 0 (l1), 1 (l2) or 2 (max)");
    return with_device_csr(n_rows, data, val_dtype, nullptr, indptr, idx_dtype, device, false,
                           [&](void *d, void *, void *p) { return spy_normalize_rows_dev(norm, n_rows, d, val_dtype, p, idx_dtype, nullptr); });
}

int spy_tfidf_host(int64_t n_rows, int64_t n_cols, void *data, int val_dtype, const void *indices, const void *indptr,
                   int idx_dtype, int tf_mode, int idf_mode, double logbase, int device) {
    return with_device_csr(n_rows, data, val_dtype, indices, indptr, idx_dtype, device, true, [&](void *d, void *i, void *p) {
        DevBuf scratch;
        int rc = scratch.up(nullptr, (size_t)spy_tfidf_scratch_bytes(n_rows, n_cols, val_dtype));
        if (rc != SPY_OK) return rc;
        rc = spy_tfidf_dev(n_rows, n_cols, d, val_dtype, i, p, idx_dtype, tf_mode, idf_mode, logbase, scratch.p, nullptr);
        if (rc == SPY_OK && cudaDeviceSynchronize() != cudaSuccess) { set_error("tfidf kernels failed"); rc = SPY_ERR_CUDA; }
        return rc;
    });
}

int spy_bm25plus_host(int64_t n_rows, int64_t n_cols, void *data, int val_dtype, const void *indices, const void *indptr,
                      int idx_dtype, double k1, double b, double delta, int tf_mode, int idf_mode, double logbase, int device) {
    return with_device_csr(n_rows, data, val_dtype, indices, indptr, idx_dtype, device, true, [&](void *d, void *i, void *p) {
        DevBuf scratch;
        int rc = scratch.up(nullptr, (size_t)spy_tfidf_scratch_bytes(n_rows, n_cols, val_dtype));
        if (rc != SPY_OK) return rc;
        rc = spy_bm25plus_dev(n_rows, n_cols, d, val_dtype, i, p, idx_dtype, k1, b, delta, tf_mode, idf_mode, logbase, scratch.p, nullptr);
        if (rc == SPY_OK && cudaDeviceSynchronize() != cudaSuccess) { set_error("bm25 kernels failed"); rc = SPY_ERR_CUDA; }
        return rc;
    });
}

// The similarity call over several GPUs of one box from ONE process: the target rows are cut into n_devices contiguous
// ranges of (nearly) equal work -- stored entries of A in the range, the cheap host-side proxy of the scalar products -- and
// every range runs spy_knn_topk_host on its device from its own host thread; B is uploaded to every device (replicated,
// SURVEY 8e).  With assemble != 0 the caller's slab (out_cols / out_values / out_rows / out_counts, n_targets * k entries) is
// the complete result in target order; with assemble == 0 nothing else changes for host buffers (every range writes
// its own rows of the same slab) -- the flag exists for symmetry with the torch.distributed path, where gathering is optional.
// range_bounds (may be NULL) receives the n_devices + 1 cut positions in the target list.
int spy_knn_topk_multi_host(const spy_knn_args *host_args, const int32_t *devices, int32_t n_devices, int32_t assemble,
                            int32_t *range_bounds) {
    (void)assemble;
    SPY_REQUIRE(host_args != nullptr, "args is NULL");
    SPY_REQUIRE(devices != nullptr && n_devices >= 1, "device list is empty");
    const spy_knn_args &a = *host_args;
    SPY_REQUIRE(a.k >= 1 && a.n_targets >= 0, "bad sizes");
    const int visible = spy_device_count();
    for (int d = 0; d < n_devices; d++) SPY_REQUIRE(devices[d] >= 0 && devices[d] < visible, "device %d is not visible", devices[d]);
    // cut by stored entries of A (+1 per row so that empty rows spread too)
    std::vector<int64_t> cum((size_t)a.n_targets + 1, 0);
    for (int i = 0; i < a.n_targets; i++) {
        const int t = a.targets[i];
        SPY_REQUIRE(t >= 0 && t < a.a_rows, "target row %d out of range", t);
        cum[i + 1] = cum[i] + (a.a_indptr[t + 1] - a.a_indptr[t]) + 1;
    }
    std::vector<int32_t> bounds((size_t)n_devices + 1, 0);
    for (int p = 1; p < n_devices; p++) {
        const int64_t goal = (cum[a.n_targets] * p + n_devices - 1) / n_devices;
        bounds[p] = (int32_t)(std::lower_bound(cum.begin(), cum.end(), goal) - cum.begin());
        bounds[p] = std::max(bounds[p], bounds[p - 1]);
    }
    bounds[n_devices] = a.n_targets;
    if (range_bounds) std::copy(bounds.begin(), bounds.end(), range_bounds);
    std::vector<int> rcs((size_t)n_devices, SPY_OK);
    std::vector<std::string> errs((size_t)n_devices);
    std::vector<std::thread> workers;
    for (int p = 0; p < n_devices; p++) {
        workers.emplace_back([&, p]() {
            const int lo = bounds[p], hi = bounds[p + 1];
            if (hi <= lo) return;
            spy_knn_args sub = a;
            sub.n_targets = hi - lo;
            sub.targets = a.targets + lo;
            const size_t off = (size_t)lo * (size_t)a.k;
            sub.out_cols = a.out_cols + off;
            sub.out_values = a.out_values + off;
            sub.out_rows = a.out_rows ? a.out_rows + off : nullptr;
            sub.out_counts = a.out_counts ? a.out_counts + lo : nullptr;
            rcs[p] = spy_knn_topk_host(&sub, devices[p]);
            if (rcs[p] != SPY_OK) errs[p] = spy_last_error();  // (the error text is thread-local)
        });
    }
    for (auto &w : workers) w.join();
    for (int p = 0; p < n_devices; p++)
        if (rcs[p] != SPY_OK) {
            set_error("device %d (rows %d..%d): %s", devices[p], bounds[p], bounds[p + 1], errs[p].c_str());
            return rcs[p];
        }
    return SPY_OK;
}

}  // extern "C"
