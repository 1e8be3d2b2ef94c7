// Micro-benchmark + semantics check: tensor memory (TMEM) as the snapshot buffer of the shared-memory accumulator panel.
//
// The hot kernel's drain (read every slot of the panel, reset it) serialises with the expansion.  Idea measured here:
// when a panel is complete, 28 warps copy it shared memory -> registers -> TMEM (LDS.128 / STS.128 sentinel /
// tcgen05.st.32x32b.x4) and clear it in the same pass; the expansion of the next panel restarts immediately while 4 other
// warps read the snapshot back with tcgen05.ld at their leisure.
//   mapping: panel column c, quad q = c / 4, tile T = q / 128 (512 columns), TMEM lane = q % 128, TMEM column = 4 T + c % 4
//   a warp can only touch lane quarter (warp id % 4): snapshot warp w handles quarter w % 4 of the tiles T = (w-4)/4 + 7 i
// Reports: cycles per snapshot pass (W floats), cycles per read-back sweep by 4 warps, and the number of mismatches.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a profiles/microbench/tmem_snapshot_bench.cu -o profiles/microbench/tmem_snapshot_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int NT = 1024;
constexpr unsigned kSent = 0x80000000u;

__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_st4(unsigned taddr, float4 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v.x)),
                 "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(unsigned taddr, unsigned (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr) : "memory");
}

// value stored in panel column c at round `it`
__device__ __forceinline__ float pattern(int c, int it) { return (float)(c + 1) + 0.25f * (float)(it & 3); }

__global__ void __launch_bounds__(NT, 1) snap(int W, int iters, unsigned long long *cycles, int *mismatch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *acc = reinterpret_cast<float *>(smem_raw);
    const unsigned acc32 = (unsigned)__cvta_generic_to_shared(acc);
    __shared__ unsigned s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nT = W / 512;  // W is a multiple of 512
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;
    const unsigned quarter = (unsigned)(warp & 3);
    const unsigned tq = tmem + ((quarter * 32u) << 16);
    const float sent = __uint_as_float(kSent);
    const float4 sent4 = make_float4(sent, sent, sent, sent);
    long long t_snap = 0, t_read = 0;
    int bad = 0;
    for (int it = 0; it < iters; it++) {
        for (int c = tid; c < W; c += NT) acc[c] = pattern(c, it);  // stands for the expansion
        __syncthreads();
        long long t0 = clock64();
        if (warp >= 4) {  // snapshot + clear by 28 warps: 7 per lane quarter
            const int m = (warp - 4) >> 2;
            for (int T = m; T < nT; T += 7) {
                const unsigned a = acc32 + (unsigned)(512 * T + 128 * (int)quarter + 4 * lane) * 4u;
                const float4 x = lds128(a);
                sts128(a, sent4);
                tmem_st4(tq + (unsigned)(4 * T), x);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        long long t1 = clock64();
        if (warp < 4) {  // read-back by 4 warps (one per lane quarter), 4 tiles per load
            for (int T = 0; T < nT; T += 4) {
                if (T + 4 <= nT) {
                    unsigned r[16];
                    tmem_ld16(tq + (unsigned)(4 * T), r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int c = 512 * (T + (i >> 2)) + 128 * (int)quarter + 4 * lane + (i & 3);
                        if (__uint_as_float(r[i]) != pattern(c, it)) bad++;
                    }
                } else {
                    for (int TT = T; TT < nT; TT++) {
                        unsigned r[4];
                        tmem_ld4(tq + (unsigned)(4 * TT), r);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int c = 512 * TT + 128 * (int)quarter + 4 * lane + i;
                            if (__uint_as_float(r[i]) != pattern(c, it)) bad++;
                        }
                    }
                }
            }
        }
        long long t2 = clock64();
        // the panel must be all sentinels now
        for (int c = tid; c < W; c += NT) if (__float_as_uint(acc[c]) != kSent) bad++;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (it > 0) { t_snap += t1 - t0; t_read += t2 - t1; }
    }
    if (bad) atomicAdd(mismatch, bad);
    if (tid == 0) { atomicAdd(cycles, (unsigned long long)t_snap); }
    if (tid == 0) { atomicAdd(cycles + 1, (unsigned long long)t_read); }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    unsigned long long *cycles; int *mismatch;
    CK(cudaMalloc(&cycles, 16)); CK(cudaMalloc(&mismatch, 4));
    for (int W : {40448, 50176, 20480}) {
        const int iters = 33;
        CK(cudaMemset(cycles, 0, 16)); CK(cudaMemset(mismatch, 0, 4));
        CK(cudaFuncSetAttribute(snap, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 4));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        CK(cudaEventRecord(e0));
        snap<<<148, NT, W * 4>>>(W, iters, cycles, mismatch);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long h[2]; int bad;
        CK(cudaMemcpy(h, cycles, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&bad, mismatch, 4, cudaMemcpyDeviceToHost));
        const double n = 148.0 * (iters - 1);
        printf("W=%6d floats (%3d KB): snapshot+clear pass %7.0f cycles, read-back sweep by 4 warps %7.0f cycles, mismatches %d (kernel %.2f ms)\n",
               W, W * 4 / 1024, h[0] / n, h[1] / n, bad, ms);
    }
    return 0;
}
