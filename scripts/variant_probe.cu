// Fast A/B probe of builds of the hot kernel through the C ABI, without Python: generates a configs[1]-shaped
// workload on the host (cosine, B = URM rows x cols with `per_row` entries per row, A shaped like URM^T),
// uploads it once, then for every library given on the command line: dlopen, plan, pack, split, three timed
// spy_knn_topk_dev launches (CUDA events), download, and a parity check of its slab against the FIRST library's.
//   nvcc -O3 -std=c++17 -Xcompiler -fopenmp -I include scripts/variant_probe.cu -o scripts/variant_probe -ldl
//   scripts/variant_probe 1000000 200000 200 100 50000 similaripy_b200/libsimilaripy_b200.so similaripy_b200/libspy_*.so
// (rows cols per_row k n_targets libs...).  A development tool: the numbers that count come from bench.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <vector>
#include <cuda_runtime.h>
#include "similaripy_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

static inline uint64_t mix(uint64_t x) { x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31); }

template <typename T> static T *upload(const std::vector<T> &v) {
    T *d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

struct Lib {
    void *h;
    int (*plan)(spy_knn_args *, int);
    int64_t (*scratch_bytes)(const spy_knn_args *, int);
    int (*build_split)(int32_t, const int32_t *, const int32_t *, int32_t, int32_t, int32_t, int32_t *, void *);
    int (*pack)(int64_t, const int32_t *, const float *, void *, void *);
    int (*topk)(const spy_knn_args *, void *, int64_t, void *);
    const char *(*last_error)(void);
    // stream engine tables (absent in round-1 builds)
    int (*chunk_counts)(int32_t, const int32_t *, const int32_t *, int32_t, int32_t, int32_t *, void *);
    int (*pad_chunks)(int32_t, const int32_t *, const int32_t *, const float *, const int32_t *, int32_t, int32_t, const int32_t *, void *, void *, int32_t);  // (ABI v5 builds ignore the trailing panel width)
    int (*row_lengths)(int32_t, const int32_t *, const int32_t *, int32_t *, void *);
    int (*build_aexp)(const spy_knn_args *, void *);
    int64_t (*scan_tmp)(int64_t);
    int (*scan32)(int64_t, const int32_t *, int32_t *, void *, void *);
    int (*scan64)(int64_t, const int32_t *, int64_t *, void *, void *);
};

int main(int argc, char **argv) {
    if (argc < 7) { printf("usage: %s rows cols per_row k n_targets lib.so [lib.so ...]\n", argv[0]); return 2; }
    const int R = atoi(argv[1]), C = atoi(argv[2]), per_row = atoi(argv[3]), k = atoi(argv[4]);
    int n_t = atoi(argv[5]);
    // ---- URM (R x C): per_row distinct sorted columns per row, values in (0.5, 1.5) ----
    std::vector<int32_t> b_indptr(R + 1), b_indices((size_t)R * per_row);
    std::vector<float> b_data((size_t)R * per_row);
    std::vector<int32_t> row_n(R);
#pragma omp parallel for schedule(static)
    for (int r = 0; r < R; r++) {
        int32_t *c = &b_indices[(size_t)r * per_row];
        for (int j = 0; j < per_row; j++) c[j] = (int32_t)(mix(((uint64_t)r << 20) + j) % (uint64_t)C);
        std::sort(c, c + per_row);
        row_n[r] = (int32_t)(std::unique(c, c + per_row) - c);
    }
    b_indptr[0] = 0;
    for (int r = 0; r < R; r++) b_indptr[r + 1] = b_indptr[r] + row_n[r];
    const int64_t nnz = b_indptr[R];
    {   // compact the rows (duplicates removed) and draw the values
        std::vector<int32_t> ci(nnz);
#pragma omp parallel for schedule(static)
        for (int r = 0; r < R; r++) memcpy(&ci[b_indptr[r]], &b_indices[(size_t)r * per_row], (size_t)row_n[r] * 4);
        b_indices.swap(ci);
        b_data.resize(nnz);
#pragma omp parallel for schedule(static)
        for (int64_t q = 0; q < nnz; q++) b_data[q] = 0.5f + (float)(mix(0xabcdefull + q) >> 40) * (1.0f / 16777216.0f);
    }
    // ---- A (C x R): rows of R * per_row / C random entries -- the shape of URM^T; it need not BE the transpose for a
    // timing + library-against-library parity probe, and generating it directly avoids a serial 2e8-entry scatter ----
    const int a_per_row = (int)((int64_t)R * per_row / C);
    std::vector<int32_t> a_indptr(C + 1), a_indices((size_t)C * a_per_row);
    std::vector<float> a_data((size_t)C * a_per_row), norm(C);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; c++) {
        a_indptr[c] = c * a_per_row;
        double sq = 0;
        for (int j = 0; j < a_per_row; j++) {
            const uint64_t h = mix(((uint64_t)c << 24) + j + 0x5555ull);
            const size_t q = (size_t)c * a_per_row + j;
            a_indices[q] = (int32_t)(h % (uint64_t)R);
            a_data[q] = 0.5f + (float)(mix(h) >> 40) * (1.0f / 16777216.0f);
            sq += (double)a_data[q] * a_data[q];
        }
        norm[c] = (float)sqrt(sq);
    }
    a_indptr[C] = C * a_per_row;
    n_t = std::min(n_t, C);
    std::vector<int32_t> targets(n_t);
    for (int i = 0; i < n_t; i++) targets[i] = (int32_t)((int64_t)i * C / n_t);
    double products = 0;
    for (int i = 0; i < n_t; i++)
        for (int32_t q = a_indptr[targets[i]]; q < a_indptr[targets[i] + 1]; q++) products += row_n[a_indices[q]];
    printf("URM %d x %d, nnz %lld; %d target rows, k=%d, %.4g products\n", R, C, (long long)nnz, n_t, k, products);
    fflush(stdout);

    // ---- device copies, shared by all libraries ----
    spy_knn_args base;
    memset(&base, 0, sizeof(base));
    base.n_targets = n_t; base.targets = upload(targets);
    base.a_rows = C; base.a_indptr = upload(a_indptr); base.a_indices = upload(a_indices); base.a_data = upload(a_data);
    base.b_rows = R; base.n_cols = C; base.b_indptr = upload(b_indptr); base.b_indices = upload(b_indices); base.b_data = upload(b_data);
    base.Xcosine = upload(norm); base.Ycosine = base.Xcosine;
    base.a1 = 1.f; base.l2 = 1.f; base.t1 = 1.f; base.t2 = 1.f; base.k = k; base.b_nnz = nnz;
    if (getenv("SPY_PROBE_DOT")) { base.l2 = 0.f; base.Xcosine = nullptr; base.Ycosine = nullptr; }  // plain dot product (configs[4]'s kind)
    const size_t slab = (size_t)n_t * k;
    int32_t *d_cols, *d_counts; float *d_vals;
    CK(cudaMalloc(&d_cols, slab * 4)); CK(cudaMalloc(&d_vals, slab * 4)); CK(cudaMalloc(&d_counts, (size_t)n_t * 4));
    base.out_cols = d_cols; base.out_values = d_vals; base.out_counts = d_counts;
    if (!base.targets || !base.a_indices || !base.a_data || !base.b_indices || !base.b_data) { printf("upload failed\n"); return 1; }
    void *d_pairs; CK(cudaMalloc(&d_pairs, ((size_t)nnz + 1) * 8));

    std::vector<int32_t> ref_cols, ref_counts; std::vector<float> ref_vals;
    for (int li = 6; li < argc; li++) {
        Lib L;
        int arg_engine = -1;  // "lib.so@2": run this library with engine 2 (stream); "@1": flat
        if (char *at = strrchr(argv[li], '@')) { arg_engine = atoi(at + 1); *at = 0; }
        L.h = dlopen(argv[li], RTLD_NOW | RTLD_LOCAL);
        if (!L.h) { printf("%-44s dlopen failed: %s\n", argv[li], dlerror()); continue; }
        L.plan = (decltype(L.plan))dlsym(L.h, "spy_knn_plan");
        L.scratch_bytes = (decltype(L.scratch_bytes))dlsym(L.h, "spy_knn_scratch_bytes");
        L.build_split = (decltype(L.build_split))dlsym(L.h, "spy_knn_build_split_dev");
        L.pack = (decltype(L.pack))dlsym(L.h, "spy_knn_pack_pairs_dev");
        L.topk = (decltype(L.topk))dlsym(L.h, "spy_knn_topk_dev");
        L.last_error = (decltype(L.last_error))dlsym(L.h, "spy_last_error");
        if (!L.plan || !L.scratch_bytes || !L.build_split || !L.pack || !L.topk) { printf("%s: missing symbols\n", argv[li]); continue; }
        L.chunk_counts = (decltype(L.chunk_counts))dlsym(L.h, "spy_knn_chunk_counts_dev");
        L.pad_chunks = (decltype(L.pad_chunks))dlsym(L.h, "spy_knn_pad_chunks_dev");
        L.row_lengths = (decltype(L.row_lengths))dlsym(L.h, "spy_knn_row_lengths_dev");
        L.build_aexp = (decltype(L.build_aexp))dlsym(L.h, "spy_knn_build_aexp_dev");
        L.scan_tmp = (decltype(L.scan_tmp))dlsym(L.h, "spy_scan_tmp_bytes");
        L.scan32 = (decltype(L.scan32))dlsym(L.h, "spy_exclusive_scan_i32_dev");
        L.scan64 = (decltype(L.scan64))dlsym(L.h, "spy_exclusive_scan_i64_dev");
        spy_knn_args a = base;
        if (const char *t = getenv("SPY_PROBE_THREADS")) a.threads = atoi(t);
        if (const char *g = getenv("SPY_PROBE_GROUP")) a.group = atoi(g);
        if (const char *g = getenv("SPY_PROBE_ENGINE")) a.engine = atoi(g);
        if (const char *g = getenv("SPY_PROBE_WIDTH")) a.panel_width = atoi(g);
        if (arg_engine >= 0) a.engine = arg_engine;
        int rc = L.plan(&a, 0);
        if (rc) { printf("%s: plan failed: %s\n", argv[li], L.last_error()); continue; }
        int32_t *d_split = nullptr;
        if (a.n_panels > 1) {
            CK(cudaMalloc(&d_split, (size_t)R * a.split_stride * 4));
            rc = L.build_split(R, a.b_indptr, a.b_indices, a.panel_width, a.n_panels, a.split_stride, d_split, nullptr);
            a.b_split = d_split;
        }
        void *d_cnt = nullptr, *d_cptr = nullptr, *d_tmp = nullptr, *d_chunks = nullptr, *d_toff = nullptr, *d_aexp = nullptr;
        float prep_ms = 0.f;
        if (!rc && a.engine == 2) {
            if (!L.chunk_counts || !L.pad_chunks || !L.row_lengths || !L.build_aexp) { printf("%s: no stream engine\n", argv[li]); continue; }
            cudaEvent_t p0, p1; cudaEventCreate(&p0); cudaEventCreate(&p1);
            const int64_t n_seg = (int64_t)R * a.n_panels;
            const int64_t n_scan = std::max<int64_t>(std::max(R, n_t), n_seg);
            CK(cudaMalloc(&d_cnt, (size_t)n_scan * 4)); CK(cudaMalloc(&d_cptr, ((size_t)n_seg + 1) * 4));
            CK(cudaMalloc(&d_tmp, (size_t)L.scan_tmp(n_scan))); CK(cudaMalloc(&d_toff, ((size_t)n_t + 1) * 8));
            CK(cudaEventRecord(p0));
            rc = L.chunk_counts(R, a.b_indptr, a.b_split, a.split_stride, a.n_panels, (int32_t *)d_cnt, nullptr);
            if (!rc) rc = L.scan32(n_seg, (const int32_t *)d_cnt, (int32_t *)d_cptr, d_tmp, nullptr);
            int32_t n_chunks = 0;
            CK(cudaMemcpy(&n_chunks, (int32_t *)d_cptr + n_seg, 4, cudaMemcpyDeviceToHost));
            CK(cudaMalloc(&d_chunks, (size_t)std::max(n_chunks, 1) * 16));
            if (!rc) rc = L.pad_chunks(R, a.b_indptr, a.b_indices, a.b_data, a.b_split, a.split_stride, a.n_panels, (const int32_t *)d_cptr, d_chunks, nullptr, a.panel_width);
            if (!rc) rc = L.row_lengths(n_t, a.targets, a.a_indptr, (int32_t *)d_cnt, nullptr);
            if (!rc) rc = L.scan64(n_t, (const int32_t *)d_cnt, (int64_t *)d_toff, d_tmp, nullptr);
            int64_t n_entries = 0;
            CK(cudaMemcpy(&n_entries, (int64_t *)d_toff + n_t, 8, cudaMemcpyDeviceToHost));
            CK(cudaMalloc(&d_aexp, (size_t)std::max<int64_t>(n_entries, 1) * a.n_panels * 8));
            a.b_chunk_indptr = (const int32_t *)d_cptr; a.b_chunks = d_chunks; a.toff = (const int64_t *)d_toff;
            a.n_entries = n_entries; a.aexp = d_aexp;
            if (!rc) rc = L.build_aexp(&a, nullptr);
            CK(cudaEventRecord(p1)); CK(cudaEventSynchronize(p1));
            cudaEventElapsedTime(&prep_ms, p0, p1);
        } else if (!rc) {
            rc = L.pack(nnz, a.b_indices, a.b_data, d_pairs, nullptr);
            a.b_pairs = d_pairs;
        }
        const int64_t sb = L.scratch_bytes(&a, 0);
        void *d_scratch = nullptr;
        CK(cudaMalloc(&d_scratch, (size_t)std::max<int64_t>(sb, 16)));
        if (rc) { printf("%s: preparation failed: %s\n", argv[li], L.last_error()); continue; }
        CK(cudaMemset(d_cols, 0xff, slab * 4)); CK(cudaMemset(d_vals, 0xff, slab * 4));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int it = 0; it < 4 && !rc; it++) {
            CK(cudaEventRecord(e0));
            rc = L.topk(&a, d_scratch, sb, nullptr);
            CK(cudaEventRecord(e1));
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) { printf("%s: launch failed: %s\n", argv[li], cudaGetErrorString(e)); return 1; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0) best = std::min(best, ms);
        }
        if (rc) { printf("%s: run failed: %s\n", argv[li], L.last_error()); continue; }
        std::vector<int32_t> cols(slab), counts(n_t); std::vector<float> vals(slab);
        CK(cudaMemcpy(cols.data(), d_cols, slab * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(vals.data(), d_vals, slab * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(counts.data(), d_counts, (size_t)n_t * 4, cudaMemcpyDeviceToHost));
        long bad_rows = 0;
        if (ref_cols.empty()) { ref_cols = cols; ref_vals = vals; ref_counts = counts; }
        else {
            // rows are written best-first: compare values position by position (rtol 1e-5) and the column ids wherever
            // the value is clearly above the row's last kept value (ties at the boundary may legitimately differ)
            for (int i = 0; i < n_t; i++) {
                bool bad = counts[i] != ref_counts[i];
                const size_t o = (size_t)i * k;
                const int n = std::min(counts[i], ref_counts[i]);
                const float last = n ? ref_vals[o + n - 1] : 0.f;
                std::vector<int32_t> x, y;
                for (int j = 0; j < n && !bad; j++) {
                    if (fabsf(vals[o + j] - ref_vals[o + j]) > 1e-5f * fabsf(ref_vals[o + j])) bad = true;
                    if (ref_vals[o + j] > last * (1.f + 4e-5f)) x.push_back(ref_cols[o + j]);
                    if (vals[o + j] > last * (1.f + 4e-5f)) y.push_back(cols[o + j]);
                }
                if (!bad) {   // every clearly-kept column of one side appears somewhere on the other side
                    std::vector<int32_t> all_r(ref_cols.begin() + o, ref_cols.begin() + o + n), all_g(cols.begin() + o, cols.begin() + o + n);
                    std::sort(all_r.begin(), all_r.end()); std::sort(all_g.begin(), all_g.end());
                    for (int32_t c : x) if (!std::binary_search(all_g.begin(), all_g.end(), c)) bad = true;
                    for (int32_t c : y) if (!std::binary_search(all_r.begin(), all_r.end(), c)) bad = true;
                }
                bad_rows += bad;
            }
        }
        if (getenv("SPY_PROBE_PHASES") && a.engine == 2) {  // -DSPY_KS_TIMING builds: 24 cycle counters at the end of the scratch
            unsigned long long ph[32];
            CK(cudaMemcpy(ph, (char *)d_scratch + sb - 256, sizeof(ph), cudaMemcpyDeviceToHost));
            const char *names[32] = {"X snapshot (all)", "X pass body", "X end-of-pass barrier", "X snapshot copy", "X wait drain", "X setup+first issue", "X staging", "",
                                     "Q snapshot (all)", "Q pass body", "Q end-of-pass barrier", "Q snapshot copy", "Q wait drain", "Q setup", "Q staging", "",
                                     "D wait snapshot", "D sweep (all)", "D forced selections", "D evaluate/tighten", "D final select+write", "#D selections", "#D slot batches", "#D failed speculations",
                                     "D speculative path", "D slot batches", "D tcgen05.ld", "#D quads queued", "#D carried bounds tried", "#D carried bounds failed", "", ""};
            for (int i = 0; i < 32; i++) if (names[i][0] && ph[1] != 0) printf("    %-26s %12.3f %s\n", names[i], ph[i] / 148.0 / (names[i][0] == '#' ? 1e3 : 1e6), names[i][0] == '#' ? "k per CTA (count)" : "Mcycles per CTA");
        }
        printf("%-44s %8.3f ms  %7.1f Gprod/s  engine %d (tables %.2f ms) panels %d x %d  threads %d group %d   rows differing from the first library: %ld\n",
               argv[li], best, products / best / 1e6, a.engine, prep_ms, a.n_panels, a.panel_width, a.threads, a.group, bad_rows);
        fflush(stdout);
        cudaFree(d_scratch); if (d_split) cudaFree(d_split);
        cudaFree(d_cnt); cudaFree(d_cptr); cudaFree(d_tmp); cudaFree(d_chunks); cudaFree(d_toff); cudaFree(d_aexp);
    }
    return 0;
}
