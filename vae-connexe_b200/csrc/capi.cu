// Library-level entry points of libcrvae_b200.so: version, error string, launch counter, device check.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace crvae {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace crvae

extern "C" int crvae_abi_version(void) { return CRVAE_ABI_VERSION; }
extern "C" const char* crvae_last_error(void) { return crvae::g_err; }
extern "C" uint64_t crvae_launch_count(void) { return crvae::g_launches.load(); }
extern "C" void crvae_launch_count_reset(void) { crvae::g_launches.store(0); }

extern "C" int crvae_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        crvae::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return CRVAE_E_NODEVICE;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess || major != 10) {
        crvae::set_error("device %d is not compute capability 10.x (sm_100a build)", dev);
        return CRVAE_E_NODEVICE;
    }
    return 0;
}
