// Shared device/host helpers for libhsb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>

#include "hsb200.h"

namespace hsb {

// ---- host side -----------------------------------------------------------------------
void set_error(const std::string& msg);           // stores into the thread-local slot
int fail(int code, const std::string& msg);       // set_error + return code
int check_launch(const char* what);               // cudaGetLastError -> status
int device_sm_count();                            // cached per device
void note_kernel(const char* name);               // records which kernel an entry point chose (hsb_last_kernel)

#define HSB_REQUIRE(cond, code, msg)                         \
    do {                                                     \
        if (!(cond)) return ::hsb::fail((code), (msg));      \
    } while (0)

struct WStrides {  // element strides of a per-patch weight tensor
    int64_t b;     // between images
    int64_t p;     // between patches of one image (row-major over fh,fw)
    int64_t k;     // between consecutive weights of one patch
};

inline WStrides make_wstrides(int w_layout, int64_t hp, int64_t P, int64_t row_stride) {
    if (w_layout == HSB_W_NCHW) return {hp * P, 1, P};
    return {P * row_stride, row_stride, 1};
}

// ---- device side ---------------------------------------------------------------------
__device__ __forceinline__ float ld_f(const float* p) { return *p; }
__device__ __forceinline__ float ld_f(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_f(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_f(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == HSB_ACT_RELU) return fmaxf(v, 0.f);
    if (act == HSB_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    return v;
}

// index of the source sample for a coordinate that may fall outside [0, n)
__device__ __forceinline__ int pad_index(int i, int n, int mode, bool& valid) {
    valid = true;
    if (i >= 0 && i < n) return i;
    switch (mode) {
        case HSB_PAD_REFLECT: {           // mirror without repeating the border sample
            if (n == 1) return 0;
            int period = 2 * (n - 1);
            int m = i % period;
            if (m < 0) m += period;
            return m < n ? m : period - m;
        }
        case HSB_PAD_REPLICATE: return i < 0 ? 0 : n - 1;
        case HSB_PAD_CIRCULAR: { int m = i % n; return m < 0 ? m + n : m; }
        default: valid = false; return 0;  // zeros
    }
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace hsb
