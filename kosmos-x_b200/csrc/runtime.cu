// Library-level plumbing: error text, launch counter, device / driver discovery.
#include "kx_internal.h"

#include <atomic>
#include <mutex>

namespace kx {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(static_cast<unsigned long long>(n), std::memory_order_relaxed); }

int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device (%s); libkosmosx_sm100 has no CPU fallback", cudaGetErrorString(e));
        return -1;
    }
    if (dev < 64 && cached[dev] > 0) return cached[dev];
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return -1;
    }
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, p.major, p.minor);
        return -1;
    }
    if (dev < 64) cached[dev] = p.multiProcessorCount;
    return p.multiProcessorCount;
}

const DriverApi& driver_api() {
    static DriverApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            api.encode_tiled = reinterpret_cast<encode_tiled_fn>(fn);
    });
    return api;
}

}  // namespace kx

extern "C" const char* kx_last_error(void) { return kx::g_err; }
extern "C" int kx_abi_version(void) { return KX_ABI_VERSION; }
extern "C" int kx_device_check(void) { return kx::device_sm_count() > 0 ? KX_OK : KX_ERR_NO_DEVICE; }
extern "C" unsigned long long kx_launch_count(void) { return kx::g_launches.load(); }
