// libhsb200: error slot, device queries and argument validation shared by every entry point.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace hsb {

static thread_local std::string g_last_error;
static thread_local const char* g_last_kernel = "";

void note_kernel(const char* name) { g_last_kernel = name; }

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const std::string& msg) {
    set_error(msg);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        return fail(HSB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    }
    return HSB_OK;
}

int device_sm_count() {
    static std::mutex mu;
    static std::vector<int> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    std::lock_guard<std::mutex> lock(mu);
    if ((int)cache.size() <= dev) cache.resize(dev + 1, 0);
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
        cache[dev] = n;
    }
    return cache[dev];
}

}  // namespace hsb

extern "C" {

int hsb_version(void) { return HSB200_VERSION; }

const char* hsb_last_error(void) { return hsb::g_last_error.c_str(); }

const char* hsb_last_kernel(void) { return hsb::g_last_kernel; }

int hsb_device_info(int* sm_count, int* compute_capability) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return hsb::fail(HSB_ERR_NO_DEVICE, "no CUDA device visible (libhsb200 has no CPU path)");
    }
    int dev = 0, major = 0, minor = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count) *sm_count = sms;
    if (compute_capability) *compute_capability = major * 10 + minor;
    return HSB_OK;
}

}  // extern "C"
