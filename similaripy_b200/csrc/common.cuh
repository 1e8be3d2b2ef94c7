// common.cuh -- shared helpers for the similaripy_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/similaripy_b200.h"

namespace spy {

// thread-local error text returned by spy_last_error()
char *err_buf();
void set_error(const char *fmt, ...);
// per-thread kernel launch counter (spy_launch_count)
void count_launch(int n = 1);

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

#define SPY_CUDA_OK(expr)                                                                       \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            spy::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                \
                           cudaGetErrorString(_e));                                             \
            return (_e == cudaErrorMemoryAllocation) ? SPY_ERR_NOMEM : SPY_ERR_CUDA;            \
        }                                                                                       \
    } while (0)

#define SPY_LAUNCH_OK()                                                                         \
    do {                                                                                        \
        spy::count_launch();                                                                    \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) {                                                                \
            spy::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,            \
                           cudaGetErrorString(_e));                                             \
            return SPY_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define SPY_REQUIRE(cond, ...)                                                                  \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            spy::set_error(__VA_ARGS__);                                                        \
            return SPY_ERR_INVALID;                                                             \
        }                                                                                       \
    } while (0)

// B200: 148 SMs, 227 KB opt-in shared memory per CTA.  Used when no device is visible
// (planning on a CPU-only host) and as the grid-sizing unit.
constexpr int kB200SmCount = 148;
constexpr int kB200MaxSmemOptin = 232448;

struct DeviceInfo {
    int sm_count;
    int max_smem_optin;
};
// device < 0 => B200 defaults without touching the CUDA runtime
DeviceInfo device_info(int device);

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace spy
