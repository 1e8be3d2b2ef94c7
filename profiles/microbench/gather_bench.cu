// Micro-benchmark: random SEGMENT gather from HBM, the access pattern of the expansion step: groups of G lanes read
// runs of `len` consecutive 8-byte pairs starting at pseudo-random (8-byte aligned) positions of a 1.6 GB array,
// 16 bytes per lane per load.  Reports algorithmic GB/s (8 bytes per pair actually used) against the sequential copy peak.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int G, int U, int PF>
__global__ void __launch_bounds__(1024, 1) gather(const uint2 *__restrict__ pairs, long n_pairs, int len, int segs_per_group, float *out) {
    const int tid = threadIdx.x, gl = tid & (G - 1);
    const unsigned gid = (blockIdx.x * 1024u + tid) / G;
    float acc = 0.f;
    for (int it = 0; it < segs_per_group; it++) {
        const long s = (long)(hash32(gid * 7919u + it) % (unsigned)(n_pairs - len - 64));
        if (PF < 0) {  // next segment into L2 by ONE bulk prefetch of exactly its bytes (cp.async.bulk.prefetch.L2)
            const long s2 = (long)(hash32(gid * 7919u + it - PF) % (unsigned)(n_pairs - len - 64));
            if (gl == 0) {
                const unsigned bytes = (unsigned)(((s2 + len - (s2 & ~1L)) * 8 + 15) & ~15L);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pairs + (s2 & ~1L)), "r"(bytes));
            }
        } else if (PF) {  // next segment into L2, lane l takes line l
            const long s2 = (long)(hash32(gid * 7919u + it + PF) % (unsigned)(n_pairs - len - 64));
            const char *nb = reinterpret_cast<const char *>(pairs + s2) + 128 * gl;
            if (nb < reinterpret_cast<const char *>(pairs + s2 + len)) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb));
        }
        const long sa = s & ~1L, e = s + len;
        for (long b = sa; b < e; b += 2 * G * U) {
            uint4 pr[U];
#pragma unroll
            for (int r = 0; r < U; r++) {
                const long q = b + 2 * gl + 2 * G * r;
                pr[r] = make_uint4(0, 0, 0, 0);
                if (q < e) pr[r] = __ldg(reinterpret_cast<const uint4 *>(pairs + q));
            }
#pragma unroll
            for (int r = 0; r < U; r++) acc += __uint_as_float(pr[r].y) + __uint_as_float(pr[r].w) + (float)(pr[r].x ^ pr[r].z);
        }
    }
    if (acc == 1.2345e-30f) out[0] = acc;
}

template <int G, int U, int PF>
void run(const char *name, const uint2 *pairs, long n_pairs, int len, float *out) {
    const int segs = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 3; it++) {
        CK(cudaEventRecord(e0));
        gather<G, U, PF><<<148, 1024>>>(pairs, n_pairs, len, segs, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double groups = 148.0 * 1024 / G, bytes = groups * segs * len * 8.0;
    printf("%-44s len=%3d  %8.3f ms  %7.1f GB/s useful  (%5.1f Gpairs/s)\n", name, len, best, bytes / best / 1e6, bytes / 8 / best / 1e6);
    fflush(stdout);
}

int main() {
    const long n_pairs = 200L * 1000 * 1000;  // 1.6 GB, like cfg2's B
    uint2 *pairs; float *out;
    CK(cudaMalloc(&pairs, n_pairs * 8 + 1024)); CK(cudaMalloc(&out, 64));
    CK(cudaMemset(pairs, 0x3c, n_pairs * 8 + 1024));
    for (int len : {50, 25, 100, 200}) {
        run<8, 2, 0>("G=8 U=2 (kernel's shape), no prefetch", pairs, n_pairs, len, out);
        run<8, 2, 1>("G=8 U=2, L2 prefetch 1 segment ahead", pairs, n_pairs, len, out);
        run<8, 2, 2>("G=8 U=2, L2 prefetch 2 segments ahead", pairs, n_pairs, len, out);
        run<8, 2, -1>("G=8 U=2, bulk L2 prefetch 1 segment ahead", pairs, n_pairs, len, out);
        run<8, 2, -2>("G=8 U=2, bulk L2 prefetch 2 segments ahead", pairs, n_pairs, len, out);
        run<8, 4, 0>("G=8 U=4, no prefetch", pairs, n_pairs, len, out);
        run<8, 4, 2>("G=8 U=4, L2 prefetch 2 ahead", pairs, n_pairs, len, out);
        run<16, 2, 2>("G=16 U=2, L2 prefetch 2 ahead", pairs, n_pairs, len, out);
        run<4, 4, 2>("G=4 U=4, L2 prefetch 2 ahead", pairs, n_pairs, len, out);
    }
    return 0;
}
