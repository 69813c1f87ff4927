// Probe for round 2: can a 5-D tensor map land an NCHW halo tile *directly* in the
// MN-major no-swizzle UMMA operand layout, so that the fused MetaBlock kernel needs no x re-stage?
//
//   global x (B, C, H, W) bf16 viewed as 5-D (innermost first): d0 = 8 pixels (16 bytes), d1 = channel (stride H*W*2),
//   d2 = 8-pixel chunk of the row (stride 16 B), d3 = row (stride W*2), d4 = image.  Box (8, KPAD, 2, TH, 1) =>
//   shared memory [row][chunk][channel][8 px]: 16-byte unit = 8 consecutive pixels of one channel, channels 16 B apart
//   (k % 8 -> 16 B, k / 8 -> LBO = 128 B), 8-pixel M-chunks KPAD*16 B apart (SBO).  Channels >= C and rows outside the
//   image are TMA zero fill.  M index = row * 16 + column of the 16-pixel-wide body of the tile.
//
// The probe loads one tile, multiplies it by a K-major weight operand with tcgen05.mma (A MN-major, B K-major),
// reads the accumulators back and compares with a host loop.  It also loads the two single-chunk halo boxes.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I hyperseg_b200/csrc -I include \
//        scripts/probe/x5d_probe.cu -o gpurun_out/x5d_probe -lcuda && gpurun_out/x5d_probe
//
// Compiled (not yet run) in round 1: no GPU minutes were left.  Expected output: "max |err| ... OK".
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../hyperseg_b200/csrc/tcgen05.cuh"
using namespace hsb;

constexpr int C = 34, H = 48, W = 64, KPAD = 48, N = 80, HID = 68;
constexpr int TH = 18, PW = 16, CHUNKS = PW / 8;                 // body of a 16x16 patch + 1 halo row above / below
constexpr int M = TH * PW;                                       // 288 body pixels
constexpr int M_TILES = (M + 127) / 128;
constexpr int A_SBO = KPAD * 16, A_LBO = 128;                    // bytes
constexpr int A_BYTES = TH * CHUNKS * A_SBO;                     // 27 648
constexpr int HALO_BYTES = TH * A_SBO;                           // one single-chunk box
constexpr int B_LBO = (N / 8) * 128, B_SBO = 128, B_BYTES = (KPAD / 8) * B_LBO;
constexpr int OFF_A = 0, OFF_PAD = OFF_A + A_BYTES;              // zero padding read by the last M tile (rows >= 288)
constexpr int PAD_BYTES = M_TILES * 128 / 8 * A_SBO - A_BYTES;
constexpr int OFF_B = OFF_PAD + PAD_BYTES, OFF_HL = OFF_B + B_BYTES, OFF_HR = OFF_HL + HALO_BYTES;
constexpr int OFF_BAR = OFF_HR + HALO_BYTES, SMEM = OFF_BAR + 64 + 1024;

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap body_map, const __grid_constant__ CUtensorMap halo_map, const __nv_bfloat16* b_packed,
      float* out, __nv_bfloat16* halo_out, int pi, int pj) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < PAD_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sm + OFF_PAD)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < B_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sm + OFF_B)[i] = reinterpret_cast<const uint4*>(b_packed)[i];
    if (tid == 0) {
        mbar_init(bar_tma, 1);
        mbar_init(bar_mma, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_tma, A_BYTES + 2 * HALO_BYTES);
        tma_load_5d(sm + OFF_A, &body_map, 0, 0, pj * CHUNKS, pi * PW - 1, 0, bar_tma);
        tma_load_5d(sm + OFF_HL, &halo_map, 0, 0, pj * CHUNKS - 1, pi * PW - 1, 0, bar_tma);      // contains column -1 (its px 7)
        tma_load_5d(sm + OFF_HR, &halo_map, 0, 0, pj * CHUNKS + CHUNKS, pi * PW - 1, 0, bar_tma); // contains column 16 (its px 0)
    }
    mbar_wait(bar_tma, 0);
    if (warp == 0 && elect_one()) {
        tc_fence_after_sync();
        constexpr uint32_t IDESC = idesc_bf16_f32(128, N, /*A MN-major*/ true, /*B K-major*/ false);
        const uint32_t a_addr = smem_u32(sm + OFF_A), b_addr = smem_u32(sm + OFF_B);
        for (int t = 0; t < M_TILES; ++t)
            for (int s = 0; s < KPAD / 16; ++s) {
                const uint64_t da = smem_desc(a_addr + t * 16 * A_SBO + 2 * s * A_LBO, A_LBO, A_SBO, SWZ_NONE);
                const uint64_t db = smem_desc(b_addr + 2 * s * B_LBO, B_LBO, B_SBO, SWZ_NONE);
                umma_bf16(tmem + t * N, da, db, IDESC, s > 0);
            }
        umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after_sync();
    for (int t = 0; t < M_TILES; ++t) {
        const int m = t * 128 + warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + t * N + c0, v);
            tmem_ld_wait();
            if (m < M)
                for (int e = 0; e < 16; ++e) out[(size_t)m * N + c0 + e] = __uint_as_float(v[e]);
        }
    }
    // halo pixels: (row r, side) -> channel c value, read the way the kernel's halo re-stage would
    for (int i = tid; i < 2 * TH * C; i += 128) {
        const int side = i / (TH * C), r = (i / C) % TH, c = i % C;
        const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(sm + (side ? OFF_HR : OFF_HL) + r * A_SBO + c * 16);
        halo_out[i] = src[side ? 0 : 7];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode fn\n"); return 1; }
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fn);
    std::vector<float> x((size_t)C * H * W), w1((size_t)HID * C);
    srand(1);
    for (auto& v : x) v = bf((rand() % 2001 - 1000) / 500.f);
    for (auto& v : w1) v = bf((rand() % 2001 - 1000) / 2000.f);
    std::vector<__nv_bfloat16> xh(x.size()), bp((size_t)B_BYTES / 2, __float2bfloat16_rn(0.f));
    for (size_t i = 0; i < x.size(); ++i) xh[i] = __float2bfloat16_rn(x[i]);
    for (int n = 0; n < HID; ++n)                         // K-major B: unit(n, kc) = kc*LBO + (n/8)*SBO + (n%8)*16 bytes
        for (int k = 0; k < C; ++k)
            bp[((k / 8) * B_LBO + (n / 8) * B_SBO + (n % 8) * 16) / 2 + k % 8] = __float2bfloat16_rn(w1[(size_t)n * C + k]);
    __nv_bfloat16 *dx, *db, *dhalo;
    float* dout;
    cudaMalloc(&dx, xh.size() * 2); cudaMalloc(&db, bp.size() * 2); cudaMalloc(&dout, (size_t)M * N * 4); cudaMalloc(&dhalo, 2 * TH * C * 2);
    cudaMemcpy(dx, xh.data(), xh.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bp.data(), bp.size() * 2, cudaMemcpyHostToDevice);
    const cuuint64_t dims[5] = {8, (cuuint64_t)C, (cuuint64_t)W / 8, (cuuint64_t)H, 1};
    const cuuint64_t strides[4] = {(cuuint64_t)H * W * 2, 16, (cuuint64_t)W * 2, (cuuint64_t)C * H * W * 2};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMap body, halo;
    const cuuint32_t box_body[5] = {8, KPAD, CHUNKS, TH, 1}, box_halo[5] = {8, KPAD, 1, TH, 1};
    CUresult r1 = encode(&body, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dx, dims, strides, box_body, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&halo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dx, dims, strides, box_halo, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode body %d halo %d (0 = ok)\n", (int)r1, (int)r2);
    if (r1 || r2) return 1;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    int bad_total = 0;
    for (int pi = 0; pi < H / PW; ++pi)
        for (int pj = 0; pj < W / PW; ++pj) {
            probe<<<1, 128, SMEM>>>(body, halo, db, dout, dhalo, pi, pj);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("patch (%d,%d): %s\n", pi, pj, cudaGetErrorString(e)); return 1; }
            std::vector<float> out((size_t)M * N);
            std::vector<__nv_bfloat16> hal(2 * TH * C);
            cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hal.data(), dhalo, hal.size() * 2, cudaMemcpyDeviceToHost);
            double worst = 0;
            int bad = 0;
            for (int m = 0; m < M; ++m) {
                const int gy = pi * PW - 1 + m / PW, gx = pj * PW + m % PW;
                for (int n = 0; n < HID; ++n) {
                    double ref = 0;
                    if (gy >= 0 && gy < H)
                        for (int c = 0; c < C; ++c) ref += (double)x[((size_t)c * H + gy) * W + gx] * w1[(size_t)n * C + c];
                    const double err = fabs(out[(size_t)m * N + n] - ref);
                    worst = err > worst ? err : worst;
                    bad += err > 1e-3 * (1 + fabs(ref));
                }
            }
            for (int i = 0; i < 2 * TH * C; ++i) {
                const int side = i / (TH * C), r = (i / C) % TH, c = i % C;
                const int gy = pi * PW - 1 + r, gx = side ? pj * PW + PW : pj * PW - 1;
                const float ref = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? x[((size_t)c * H + gy) * W + gx] : 0.f;
                bad += fabs(__bfloat162float(hal[i]) - ref) > 0;
            }
            printf("patch (%d,%d): max |err| %.3g, mismatches %d %s\n", pi, pj, worst, bad, bad ? "FAIL" : "OK");
            bad_total += bad;
        }
    printf(bad_total ? "x5d probe FAILED\n" : "x5d probe OK: TMA -> MN-major UMMA operand works without a re-stage\n");
    return bad_total != 0;
}
