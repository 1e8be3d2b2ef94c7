// api.cu -- library-level entry points: version, device discovery, error text, launch counter.
#include "common.cuh"
#include <cstring>

namespace spy {

static thread_local char g_err[1024] = "";
static thread_local long long g_launches = 0;

char *err_buf() { return g_err; }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += n; }

DeviceInfo device_info(int device) {
    DeviceInfo d{kB200SmCount, kB200MaxSmemOptin};
    if (device < 0) return d;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) d.sm_count = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess && v > 0)
        d.max_smem_optin = v;
    cudaGetLastError();
    return d;
}

}  // namespace spy

extern "C" {

int spy_abi_version(void) { return SPY_ABI_VERSION; }

int spy_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *spy_last_error(void) { return spy::err_buf(); }

int64_t spy_launch_count(int reset) {
    long long v = spy::g_launches;
    if (reset) spy::g_launches = 0;
    return v;
}

}  // extern "C"
