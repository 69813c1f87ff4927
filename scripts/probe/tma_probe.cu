// Probe: which tiled-TMA configurations load correctly on this GPU (isolates UTMALDG problems).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hyperseg_b200/csrc/tcgen05.cuh"
using namespace hsb;

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap map, __nv_bfloat16* out, int n, int c0, int c1, int c2, int c3) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(sm + 1024);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, n * 2);
        if (RANK == 4) tma_load_4d(dst, &map, c0, c1, c2, c3, bar);
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(dst)), "l"(&map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(dst)), "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int only = argc > 1 ? atoi(argv[1]) : -1;
    int idx = -1;
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    EncodeTiledFn encode = (EncodeTiledFn)ptr;
    printf("encode fn %p q=%d\n", ptr, (int)q);
    const int W = 64, H = 48, C = 34, B = 2;
    size_t n = (size_t)W * H * C * B;
    std::vector<__nv_bfloat16> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16((float)(i % 251));
    __nv_bfloat16 *d, *out;
    cudaMalloc(&d, n * 2); cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 65536 * 2);
    struct Cfg { int rank; int box[4]; int c[4]; const char* name; };
    Cfg cfgs[] = {
        {2, {32, 8, 1, 1}, {0, 0, 0, 0}, "2d box32x8 @0,0"},
        {2, {32, 8, 1, 1}, {15, 15, 0, 0}, "2d box32x8 @15,15"},
        {2, {32, 8, 1, 1}, {16, 15, 0, 0}, "2d box32x8 @16,15"},
        {2, {32, 8, 1, 1}, {-1, -1, 0, 0}, "2d box32x8 @-1,-1"},
        {2, {24, 8, 1, 1}, {0, 0, 0, 0}, "2d box24x8 @0,0"},
        {2, {32, 18, 1, 1}, {0, 0, 0, 0}, "2d box32x18 @0,0"},
        {2, {24, 18, 1, 1}, {15, 15, 0, 0}, "2d box24x18 @15,15"},
        {3, {32, 8, 4, 1}, {0, 0, 0, 0}, "3d box32x8x4 @0"},
        {4, {32, 8, 4, 1}, {0, 0, 0, 1}, "4d box32x8x4 @0,b1"},
        {4, {24, 18, 34, 1}, {15, 15, 0, 1}, "4d box24x18x34 @15,15"},
        {4, {24, 18, 34, 1}, {-1, -1, 0, 0}, "4d box24x18x34 @-1,-1"},
    };
    for (auto& cf : cfgs) {
        ++idx;
        if (only >= 0 && idx != only) continue;
        CUtensorMap map;
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)W * 2, (cuuint64_t)W * H * 2, (cuuint64_t)W * H * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)cf.box[0], (cuuint32_t)cf.box[1], (cuuint32_t)cf.box[2], (cuuint32_t)cf.box[3]};
        cuuint32_t es[4] = {1, 1, 1, 1};
        if (cf.rank == 2) { dims[1] = (cuuint64_t)H * C * B; }
        if (cf.rank == 3) { dims[2] = (cuuint64_t)C * B; }
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, cf.rank, d, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int nel = cf.box[0] * cf.box[1] * (cf.rank >= 3 ? cf.box[2] : 1) * (cf.rank >= 4 ? cf.box[3] : 1);
        cudaMemset(out, 0, 65536 * 2);
        size_t smem = 1024 + nel * 2 + 128;
        cudaError_t e = cudaSuccess;
        if (cf.rank == 2) { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<2><<<1, 128, smem>>>(map, out, nel, cf.c[0], cf.c[1], cf.c[2], cf.c[3]); }
        if (cf.rank == 3) { cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<3><<<1, 128, smem>>>(map, out, nel, cf.c[0], cf.c[1], cf.c[2], cf.c[3]); }
        if (cf.rank == 4) { cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<4><<<1, 128, smem>>>(map, out, nel, cf.c[0], cf.c[1], cf.c[2], cf.c[3]); }
        e = cudaDeviceSynchronize();
        std::vector<__nv_bfloat16> ho(nel);
        int bad = -1;
        if (e == cudaSuccess) {
            cudaMemcpy(ho.data(), out, nel * 2, cudaMemcpyDeviceToHost);
            bad = 0;
            for (int i = 0; i < nel; ++i) {
                int q0 = i % cf.box[0], q1 = (i / cf.box[0]) % cf.box[1], q2 = (i / (cf.box[0] * cf.box[1])) % cf.box[2];
                long x = cf.c[0] + q0, y = cf.c[1] + q1, c = cf.c[2] + q2, b = cf.c[3];
                float expect = 0.f;
                bool in = x >= 0 && x < W && y >= 0;
                if (cf.rank == 2) in = in && y < (long)H * C * B;
                else in = in && y < H;
                if (in) {
                    size_t idx = (cf.rank == 2) ? (size_t)y * W + x : (cf.rank == 3 ? ((size_t)c * H + y) * W + x : (((size_t)b * C + c) * H + y) * W + x);
                    expect = __bfloat162float(h[idx]);
                }
                if (__bfloat162float(ho[i]) != expect) ++bad;
            }
        }
        printf("%-28s encode=%d launch=%s mismatches=%d of %d\n", cf.name, (int)r, cudaGetErrorString(e), bad, nel);
        if (e != cudaSuccess) { printf("sticky error, stopping\n"); break; }
    }
    return 0;
}
