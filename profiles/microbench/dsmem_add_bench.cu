// Micro-benchmark: float adds into a shared-memory panel that is spread over a 2-CTA thread-block cluster
// (distributed shared memory): red.shared::cluster.add.f32 to the CTA's own panel, to the partner's panel, and to a
// 50/50 mix.  Question: would a cluster-wide accumulator (half the panels, twice the segment length per target row)
// pay for the hot kernel?  Reports Gadd/s over 148 SMs (74 clusters).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a profiles/microbench/dsmem_add_bench.cu -o profiles/microbench/dsmem_add_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int NT = 1024;
__host__ __device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// MODE 0: own panel (plain red.shared), 1: own panel through the cluster address, 2: partner's panel, 3: 50/50 mix
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) adds(int W, int iters, float *out) {
    extern __shared__ __align__(16) float acc[];
    const unsigned own = (unsigned)__cvta_generic_to_shared(acc);
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    unsigned own_c, other_c;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(own_c) : "r"(own), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(other_c) : "r"(own), "r"(rank ^ 1u));
    const int tid = threadIdx.x;
    for (int i = tid; i < W; i += NT) acc[i] = 0.f;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    unsigned h = hash32(blockIdx.x * 1024u + tid);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            h = hash32(h + 0x9e3779b9u);
            const unsigned col = h % (unsigned)W;
            if (MODE == 0) asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(own + col * 4u), "f"(1.0f) : "memory");
            else {
                const unsigned base = (MODE == 1) ? own_c : (MODE == 2) ? other_c : ((h >> 20) & 1u) ? other_c : own_c;
                asm volatile("red.shared::cluster.add.f32 [%0], %1;" ::"r"(base + col * 4u), "f"(1.0f) : "memory");
            }
        }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    float t = 0.f;
    for (int i = tid; i < W; i += NT) t += acc[i];
    atomicAdd(out, t);
}

template <int MODE>
void run(const char *name, int W, float *out) {
    const int iters = 2048;
    CK(cudaFuncSetAttribute(adds<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, h = 0.f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaMemset(out, 0, 4));
        CK(cudaEventRecord(e0));
        adds<MODE><<<148, NT, W * 4>>>(W, iters, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
        CK(cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost));
    }
    const double n = 148.0 * NT * iters * 4.0;
    printf("%-52s W=%5d  %8.3f ms  %7.1f Gadd/s   sum %s\n", name, W, best, n / best / 1e6, (double)h == n ? "ok" : "(rounded)");
}

int main() {
    float *out; CK(cudaMalloc(&out, 4));
    const int W = 50048;
    run<0>("own panel, red.shared.add.f32", W, out);
    run<1>("own panel, red.shared::cluster.add.f32", W, out);
    run<2>("partner's panel, red.shared::cluster.add.f32", W, out);
    run<3>("50/50 own / partner, red.shared::cluster.add.f32", W, out);
    return 0;
}
