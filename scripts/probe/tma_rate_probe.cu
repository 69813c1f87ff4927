// How fast does one SM's TMA unit deliver the x tile of a 16x16 patch (34 channels, 18 rows), depending on how the tile is cut
// into boxes?  Every CTA (one per SM) streams the tiles of its patches through a 2-deep ring, no compute.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I hyperseg_b200/csrc -I include \
//        scripts/probe/tma_rate_probe.cu -o gpurun_out/tma_rate_probe -lcuda && gpurun_out/tma_rate_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "../../hyperseg_b200/csrc/tcgen05.cuh"
using namespace hsb;

constexpr int CIN = 34, TH = 18, PS = 16, B = 8, H = 256, W = 512, FH = H / PS, FW = W / PS, TOTAL = B * FH * FW;
constexpr int STAGE = 64 * 1024;

struct Maps { CUtensorMap m[4]; };

// modes: see main()
template <int MODE>
__global__ void __launch_bounds__(32, 1) rate(const __grid_constant__ Maps maps, int contiguous, long long* cycles) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * STAGE);
    if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int per = (TOTAL + gridDim.x - 1) / gridDim.x;
    auto patch_of = [&](int i) { return contiguous ? blockIdx.x * per + i : blockIdx.x + i * (int)gridDim.x; };
    const int n = contiguous ? min(per, TOTAL - (int)blockIdx.x * per) : (TOTAL - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto issue = [&](int patch, int st) {
        const int pp = patch % (FH * FW), b = patch / (FH * FW), pi = pp / FW, pj = pp % FW;
        unsigned char* dst = sm + st * STAGE;
        uint64_t* bar = bars + st;
        const int y0 = pi * PS - 1, x0 = pj * PS;
        if (MODE == 0) {            // 5 ops, 5-D SW32 {16, 8, 18, 1, 1}: the kernel's body load
            mbar_arrive_expect_tx(bar, 5 * TH * 256);
            for (int kg = 0; kg < 5; ++kg) tma_load_5d(dst + kg * TH * 256, kg < 4 ? &maps.m[0] : &maps.m[1], x0, 0, y0, kg < 4 ? kg : 0, b, bar);
        } else if (MODE == 1) {     // 1 op, 4-D {16, 18, 34}: aligned 32-byte rows
            mbar_arrive_expect_tx(bar, CIN * TH * 32);
            tma_load_4d(dst, &maps.m[2], x0, y0, 0, b, bar);
        } else if (MODE == 2) {     // 1 op, 4-D {8, 18, 34}: 16-byte rows (one halo chunk)
            mbar_arrive_expect_tx(bar, CIN * TH * 16);
            tma_load_4d(dst, &maps.m[3], x0 - 8, y0, 0, b, bar);
        } else if (MODE == 3) {     // body (mode 0) + both halo chunks (mode 2 x 2)
            mbar_arrive_expect_tx(bar, 5 * TH * 256 + 2 * CIN * TH * 16);
            for (int kg = 0; kg < 5; ++kg) tma_load_5d(dst + kg * TH * 256, kg < 4 ? &maps.m[0] : &maps.m[1], x0, 0, y0, kg < 4 ? kg : 0, b, bar);
            tma_load_4d(dst + 24 * 1024, &maps.m[3], x0 - 8, y0, 0, b, bar);
            tma_load_4d(dst + 36 * 1024, &maps.m[3], x0 + PS, y0, 0, b, bar);
        } else if (MODE == 4) {     // 2 ops: 4 k-groups in one 5-D op + the tail group
            mbar_arrive_expect_tx(bar, 5 * TH * 256);
            tma_load_5d(dst, &maps.m[0], x0, 0, y0, 0, b, bar);     // box {16, 8, 18, 4, 1} (map 0 rebuilt by the host for this mode)
            tma_load_5d(dst + 4 * TH * 256, &maps.m[1], x0, 0, y0, 0, b, bar);
        }
    };
    const long long t0 = clock64();
    if (n > 0) issue(patch_of(0), 0);
    for (int it = 0; it < n; ++it) {
        if (it + 1 < n) issue(patch_of(it + 1), (it + 1) & 1);
        mbar_wait(bars + (it & 1), (it >> 1) & 1);
    }
    cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE>
static void run(const char* what, const Maps& maps, long long* dcyc) {
    cudaFuncSetAttribute(rate<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * STAGE + 2048);
    for (int contiguous = 0; contiguous < 2; ++contiguous) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            rate<MODE><<<148, 32, 2 * STAGE + 2048>>>(maps, contiguous, dcyc);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(e)); return; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        printf("mode %d %-44s %s walk: %6.1f us -> %5.0f cycles/patch/SM (at 1.965 GHz)\n", MODE, what, contiguous ? "row " : "grid", best * 1e3,
               best * 1e-3 * 1.965e9 / (TOTAL / 148.0));
    }
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode fn\n"); return 1; }
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    __nv_bfloat16* dx;
    cudaMalloc(&dx, (size_t)B * CIN * H * W * 2 + 65536);
    cudaMemset(dx, 0, (size_t)B * CIN * H * W * 2 + 65536);
    long long* dcyc;
    cudaMalloc(&dcyc, 1024 * sizeof(long long));
    const cuuint64_t HW2 = (cuuint64_t)H * W * 2;
    const cuuint32_t e5[5] = {1, 1, 1, 1, 1};
    Maps maps;
    auto make5 = [&](CUtensorMap* m, void* base, int ch_in_group, int groups, int box_groups) {
        const cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)ch_in_group, (cuuint64_t)H, (cuuint64_t)groups, (cuuint64_t)B};
        const cuuint64_t strides[4] = {HW2, (cuuint64_t)W * 2, 8 * HW2, (cuuint64_t)CIN * HW2};
        const cuuint32_t box[5] = {16, 8, (cuuint32_t)TH, (cuuint32_t)box_groups, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, e5, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    auto make4 = [&](CUtensorMap* m, int box_w) {
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)CIN, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)W * 2, HW2, (cuuint64_t)CIN * HW2};
        const cuuint32_t box[4] = {(cuuint32_t)box_w, (cuuint32_t)TH, (cuuint32_t)CIN, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, box, e5, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    int r = 0;
    r |= make5(&maps.m[0], dx, 8, 4, 1);
    r |= make5(&maps.m[1], dx + (size_t)32 * H * W, 2, 1, 1);
    r |= make4(&maps.m[2], 16);
    r |= make4(&maps.m[3], 8);
    printf("encode %d\n", r);
    run<0>("5 ops 5-D SW32 {16,8,18,1}", maps, dcyc);
    run<1>("1 op 4-D {16,18,34} 32-byte rows", maps, dcyc);
    run<2>("1 op 4-D {8,18,34} 16-byte rows", maps, dcyc);
    run<3>("5 ops body + 2 ops halo chunks", maps, dcyc);
    r = make5(&maps.m[0], dx, 8, 4, 4);
    printf("encode 4-group box %d\n", r);
    run<4>("2 ops 5-D SW32 {16,8,18,4} + tail", maps, dcyc);
    return 0;
}
