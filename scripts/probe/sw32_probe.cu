// Round-2 probe for the restage-free fused MetaBlock kernel:
//
//  (1) a 5-D tensor map over NCHW x, dims (innermost first) = {W, c%8, H, c/8, B}, box {PW, 8, TH, 1, 1} with
//      SWIZZLE_32B (PW = 16) or no swizzle (PW = 8), lands the body of a halo tile as [row][c%8][PW px] per 8-channel
//      group -- which IS the canonical MN-major UMMA operand (M-group = one tile row, K rows 32 / 16 bytes apart);
//  (2) the 2*TH halo-column pixels are written by threads into extra M-groups (manual swizzle);
//  (3) tcgen05.mma with an MN-major swizzled A and a K-major B whose rows are only HID (not N) apart (over-read);
//  (4) the BatchNorm shift enters through an "init" MMA: D = Ones(128x16, all strides 0) . ShiftB, accumulate = false;
//  (5) a TMA rate test: how fast can one elected thread per CTA stream such boxes (bytes / cycle / SM)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I hyperseg_b200/csrc -I include \
//        scripts/probe/sw32_probe.cu -o gpurun_out/sw32_probe -lcuda && gpurun_out/sw32_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../hyperseg_b200/csrc/tcgen05.cuh"
using namespace hsb;

constexpr uint32_t SWZ_32B = 6;

template <int CIN_, int HID_, int PS_>
struct Cfg {
    static constexpr int CIN = CIN_, HID = HID_, PS = PS_, TH = PS_ + 2;
    static constexpr int KG = (CIN + 7) / 8;                 // 8-channel groups that carry data
    static constexpr int K1 = (CIN + 15) / 16 * 16, N1 = (HID + 15) / 16 * 16;
    static constexpr int ROWB = PS * 2;                      // bytes of one (channel, tile row): 32 or 16
    static constexpr int GRP = 8 * ROWB;                     // one M-group of one k-group: 256 or 128 bytes
    static constexpr int HALO_G = (2 * TH + PS - 1) / PS;    // M-groups that hold the halo-column pixels
    static constexpr int MG = TH + HALO_G;                   // M-groups per k-group
    static constexpr int A_LBO = GRP, A_SBO = MG * GRP;      // swizzled MN-major: LBO = M-group stride, SBO = k-group stride
    static constexpr int M = TH * PS + 2 * TH;               // rows that matter
    static constexpr int MT = (M + 127) / 128;
    static constexpr int A_BYTES = (K1 / 8) * A_SBO + 8 * GRP;   // + over-read of the last M tile
    static constexpr int B_LBO = HID * 16, B_SBO = 128;      // arranged weights: [kc][n < HID][8 k]
    static constexpr int B_DATA = KG * B_LBO;
    static constexpr int B_BYTES = (K1 / 8 - 1) * B_LBO + N1 * 16 > B_DATA ? (K1 / 8 - 1) * B_LBO + N1 * 16 : B_DATA;
    static constexpr int SH_LBO = N1 * 16, SH_BYTES = 2 * SH_LBO;
    static constexpr int OFF_A = 0, OFF_B = (A_BYTES + 255) / 256 * 256, OFF_SH = OFF_B + (B_BYTES + 127) / 128 * 128;
    static constexpr int OFF_ONES = OFF_SH + SH_BYTES, OFF_BAR = OFF_ONES + 128, SMEM = OFF_BAR + 64 + 1024;
    static constexpr int TMEM_COLS = 256;
    static_assert(MT * N1 <= TMEM_COLS, "tmem");
};

// byte offset of element (k-group kg, M-group g, channel c8, pixel j) inside A, swizzle applied (PS = 16 only)
template <class C>
__host__ __device__ inline int a_off(int kg, int g, int c8, int j) {
    int o = kg * C::A_SBO + g * C::GRP + c8 * C::ROWB + j * 2;
    if (C::PS == 16) o ^= ((o >> 7) & 1) << 4;
    return o;
}

template <class C>
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_tail, const __nv_bfloat16* x,
      const __nv_bfloat16* b_arranged, const float* shift, float* out, int H, int W, int pi, int pj, int ones_mode) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // poison A with NaN patterns (rows nobody reads may hold anything), zero B (padding must be zero)
    for (int i = tid; i < C::OFF_B / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x7FC07FC0u;
    for (int i = tid; i < (C::OFF_BAR - C::OFF_B) / 4; i += 128) reinterpret_cast<uint32_t*>(sm + C::OFF_B)[i] = 0u;
    __syncthreads();
    for (int i = tid; i < C::B_DATA / 16; i += 128) reinterpret_cast<uint4*>(sm + C::OFF_B)[i] = reinterpret_cast<const uint4*>(b_arranged)[i];
    // zero the k-groups that the TMA does not fill (K1/8 > KG): A rows there meet finite B over-read bytes
    for (int i = tid; i < (C::K1 / 8 - C::KG) * C::A_SBO / 4; i += 128) reinterpret_cast<uint32_t*>(sm + C::KG * C::A_SBO)[i] = 0u;
    for (int n = tid; n < C::N1; n += 128) {   // ShiftB: K-major no swizzle, element (n, k = 0) = shift
        __nv_bfloat16* u = reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_SH + (n >> 3) * 128 + (n & 7) * 16);
        u[0] = __float2bfloat16_rn(n < C::HID ? shift[n] : 0.f);
    }
    for (int i = tid; i < 64; i += 128) reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_ONES)[i] = __float2bfloat16_rn(1.f);
    if (tid == 0) {
        mbar_init(bar_tma, 1);
        mbar_init(bar_mma, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, C::TMEM_COLS);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_tma, C::KG * C::TH * C::GRP);
        for (int kg = 0; kg < C::KG; ++kg) {
            const bool tail = (C::CIN % 8 != 0) && kg == C::KG - 1;
            tma_load_5d(sm + C::OFF_A + kg * C::A_SBO, tail ? &map_tail : &map_full, pj * C::PS, 0, pi * C::PS - 1, tail ? 0 : kg, 0, bar_tma);
        }
    }
    // halo-column pixels straight from global memory (reflect handled in the index), two per thread-iteration
    for (int i = tid; i < C::CIN * C::TH; i += 128) {
        const int c = i / C::TH, r = i % C::TH;
        int gy = pi * C::PS - 1 + r;
        gy = gy < 0 ? -gy : (gy >= H ? 2 * H - 2 - gy : gy);
        int gl = pj * C::PS - 1, gr = pj * C::PS + C::PS;
        gl = gl < 0 ? -gl : gl;
        gr = gr >= W ? 2 * W - 2 - gr : gr;
        const __nv_bfloat16* row = x + ((size_t)c * H + gy) * W;
        const int hl = r, hr = C::TH + r;
        *reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_A + a_off<C>(c >> 3, C::TH + hl / C::PS, c & 7, hl % C::PS)) = row[gl];
        *reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_A + a_off<C>(c >> 3, C::TH + hr / C::PS, c & 7, hr % C::PS)) = row[gr];
    }
    // channels CIN .. 8*KG-1 of the halo groups must be zero as well (the TMA zero-fills them for the body)
    for (int i = tid; i < (8 * C::KG - C::CIN) * 2 * C::TH; i += 128) {
        const int c = C::CIN + i / (2 * C::TH), h = i % (2 * C::TH);
        *reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_A + a_off<C>(c >> 3, C::TH + h / C::PS, c & 7, h % C::PS)) = __float2bfloat16_rn(0.f);
    }
    mbar_wait(bar_tma, 0);
    // reflect rows of the body at the image border: tile row 0 <- row 2, row TH-1 <- row TH-3 (whole M-groups)
    if (pi == 0 || (pi + 1) * C::PS == H) {
        for (int i = tid; i < C::KG * C::GRP / 16; i += 128) {
            const int kg = i / (C::GRP / 16), u = i % (C::GRP / 16);
            unsigned char* base = sm + C::OFF_A + kg * C::A_SBO + u * 16;
            if (pi == 0) *reinterpret_cast<uint4*>(base) = *reinterpret_cast<uint4*>(base + 2 * C::GRP);
            if ((pi + 1) * C::PS == H) *reinterpret_cast<uint4*>(base + (C::TH - 1) * C::GRP) = *reinterpret_cast<uint4*>(base + (C::TH - 3) * C::GRP);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {
        tc_fence_after_sync();
        constexpr uint32_t IDESC = idesc_bf16_f32(128, C::N1, /*A MN-major*/ true, /*B K-major*/ false);
        constexpr uint32_t IDESC_INIT = idesc_bf16_f32(128, C::N1, false, false);
        const uint32_t a_addr = smem_u32(sm + C::OFF_A), b_addr = smem_u32(sm + C::OFF_B);
        const uint32_t sh_addr = smem_u32(sm + C::OFF_SH), ones_addr = smem_u32(sm + C::OFF_ONES);
        for (int t = 0; t < C::MT; ++t) {
            if (ones_mode) umma_bf16(tmem + t * C::N1, smem_desc(ones_addr, 0, 0, SWZ_NONE), smem_desc(sh_addr, C::SH_LBO, 128, SWZ_NONE), IDESC_INIT, false);
            for (int s = 0; s < C::K1 / 16; ++s) {
                uint64_t da;
                if (C::PS == 16) da = smem_desc(a_addr + 2 * s * C::A_SBO + t * (128 / C::PS) * C::A_LBO, C::A_LBO, C::A_SBO, SWZ_32B);
                else             da = smem_desc(a_addr + 2 * s * C::A_SBO + t * (128 / C::PS) * C::A_LBO, /*k-group*/ C::A_SBO, /*M-group*/ C::A_LBO, SWZ_NONE);
                const uint64_t db = smem_desc(b_addr + 2 * s * C::B_LBO, C::B_LBO, C::B_SBO, SWZ_NONE);
                umma_bf16(tmem + t * C::N1, da, db, IDESC, ones_mode || s > 0);
            }
        }
        umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after_sync();
    for (int t = 0; t < C::MT; ++t) {
        const int m = t * 128 + warp * 32 + lane;
        for (int c0 = 0; c0 < C::N1; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + t * C::N1 + c0, v);
            tmem_ld_wait();
            if (m < C::M)
                for (int e = 0; e < 16; ++e) out[(size_t)m * C::N1 + c0 + e] = __uint_as_float(v[e]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
}

// ---- TMA rate test: every CTA streams boxes of `patches` patches through a 2-deep ring, no compute ----------------------
template <class C, int MODE>   // MODE 0: this probe's 5-D boxes; 1: round-1 4-D box {32 px, TH, CIN}
__global__ void __launch_bounds__(32, 1)
rate(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_tail, const __grid_constant__ CUtensorMap map4,
     int fh, int fw, int total, long long* cycles) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    constexpr int STAGE = MODE == 0 ? C::KG * C::TH * C::GRP : C::CIN * C::TH * 64;
    constexpr int STAGE_A = (STAGE + 1023) / 1024 * 1024;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * STAGE_A);
    if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    auto issue = [&](int patch, int st) {
        const int pp = patch % (fh * fw), b = patch / (fh * fw), pi = pp / fw, pj = pp % fw;
        mbar_arrive_expect_tx(bars + st, STAGE);
        if (MODE == 0) {
            for (int kg = 0; kg < C::KG; ++kg) {
                const bool tail = (C::CIN % 8 != 0) && kg == C::KG - 1;
                tma_load_5d(sm + st * STAGE_A + kg * C::TH * C::GRP, tail ? &map_tail : &map_full, pj * C::PS, 0, pi * C::PS - 1, tail ? 0 : kg, b, bars + st);
            }
        } else {
            tma_load_4d(sm + st * STAGE_A, &map4, pj * C::PS - 8, pi * C::PS - 1, 0, b, bars + st);
        }
    };
    int it = 0;
    if ((int)blockIdx.x < total) issue(blockIdx.x, 0);
    for (int patch = blockIdx.x; patch < total; patch += gridDim.x, ++it) {
        const int next = patch + gridDim.x;
        if (next < total) issue(next, (it + 1) & 1);
        mbar_wait(bars + (it & 1), (it >> 1) & 1);
    }
    cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode;
static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <class C>
static bool make_maps(__nv_bfloat16* dx, int B, int H, int W, CUtensorMap* full, CUtensorMap* tail) {
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const cuuint32_t box[5] = {(cuuint32_t)C::PS, 8, (cuuint32_t)C::TH, 1, 1};
    const CUtensorMapSwizzle swz = C::PS == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const cuuint64_t HW2 = (cuuint64_t)H * W * 2;
    const cuuint64_t dims[5] = {(cuuint64_t)W, 8, (cuuint64_t)H, (cuuint64_t)(C::CIN / 8 > 0 ? C::CIN / 8 : 1), (cuuint64_t)B};
    const cuuint64_t strides[4] = {HW2, (cuuint64_t)W * 2, 8 * HW2, (cuuint64_t)C::CIN * HW2};
    CUresult r1 = g_encode(full, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = CUDA_SUCCESS;
    if (C::CIN % 8) {
        const cuuint64_t dimt[5] = {(cuuint64_t)W, (cuuint64_t)(C::CIN % 8), (cuuint64_t)H, 1, (cuuint64_t)B};
        r2 = g_encode(tail, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dx + (size_t)(C::CIN / 8) * 8 * H * W, dimt, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        *tail = *full;
    }
    printf("  encode full %d tail %d (0 = ok)\n", (int)r1, (int)r2);
    return r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS;
}

template <class C>
static int run_case(const char* name) {
    constexpr int H = 3 * C::PS, W = 4 * C::PS;
    printf("== %s: CIN %d HID %d PS %d  (M %d, tiles %d, A %d B, SBO %d)\n", name, C::CIN, C::HID, C::PS, C::M, C::MT, C::A_BYTES, C::A_SBO);
    std::vector<float> x((size_t)C::CIN * H * W), w1((size_t)C::HID * C::CIN), sh(C::HID);
    srand(1);
    for (auto& v : x) v = bf((rand() % 2001 - 1000) / 500.f);
    for (auto& v : w1) v = bf((rand() % 2001 - 1000) / 2000.f);
    for (auto& v : sh) v = bf((rand() % 2001 - 1000) / 1000.f);
    std::vector<__nv_bfloat16> xh(x.size()), bp((size_t)C::B_DATA / 2, __float2bfloat16_rn(0.f));
    for (size_t i = 0; i < x.size(); ++i) xh[i] = __float2bfloat16_rn(x[i]);
    for (int n = 0; n < C::HID; ++n)
        for (int k = 0; k < C::CIN; ++k) bp[((k / 8) * C::B_LBO + n * 16) / 2 + k % 8] = __float2bfloat16_rn(w1[(size_t)n * C::CIN + k]);
    __nv_bfloat16 *dx, *db;
    float *dout, *dsh;
    cudaMalloc(&dx, xh.size() * 2); cudaMalloc(&db, bp.size() * 2); cudaMalloc(&dout, (size_t)C::M * C::N1 * 4); cudaMalloc(&dsh, C::HID * 4);
    cudaMemcpy(dx, xh.data(), xh.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bp.data(), bp.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dsh, sh.data(), C::HID * 4, cudaMemcpyHostToDevice);
    CUtensorMap full, tail;
    if (!make_maps<C>(dx, 1, H, W, &full, &tail)) return 1;
    cudaFuncSetAttribute(probe<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    int bad_total = 0;
    for (int ones_mode = 0; ones_mode < 2; ++ones_mode)
        for (int pi = 0; pi < H / C::PS; ++pi)
            for (int pj = 0; pj < W / C::PS; ++pj) {
                cudaMemset(dout, 0xFF, (size_t)C::M * C::N1 * 4);
                probe<C><<<1, 128, C::SMEM>>>(full, tail, dx, db, dsh, dout, H, W, pi, pj, ones_mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("  patch (%d,%d): %s\n", pi, pj, cudaGetErrorString(e)); return 1; }
                std::vector<float> out((size_t)C::M * C::N1);
                cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
                double worst = 0;
                int bad = 0;
                auto refl = [](int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); };
                for (int m = 0; m < C::M; ++m) {
                    int r, col;                       // tile coordinates: row 0..TH-1, column -1..PS
                    if (m < C::TH * C::PS) { r = m / C::PS; col = m % C::PS; }
                    else { const int h = m - C::TH * C::PS; r = h % C::TH; col = h < C::TH ? -1 : C::PS; }
                    const int gy = refl(pi * C::PS - 1 + r, H), gx = refl(pj * C::PS + col, W);
                    for (int n = 0; n < C::HID; ++n) {
                        double ref = ones_mode ? sh[n] : 0.0;
                        for (int c = 0; c < C::CIN; ++c) ref += (double)x[((size_t)c * H + gy) * W + gx] * w1[(size_t)n * C::CIN + c];
                        const double err = fabs(out[(size_t)m * C::N1 + n] - ref);
                        if (!(err <= 1e-3 * (1 + fabs(ref)))) ++bad;
                        if (err > worst || err != err) worst = err;
                    }
                }
                printf("  shift-mma %d patch (%d,%d): max |err| %.3g, mismatches %d %s\n", ones_mode, pi, pj, worst, bad, bad ? "FAIL" : "OK");
                bad_total += bad;
            }
    printf(bad_total ? "  %s FAILED\n" : "  %s OK\n", name);
    cudaFree(dx); cudaFree(db); cudaFree(dout); cudaFree(dsh);
    return bad_total != 0;
}

template <class C>
static void run_rate(const char* name, int B, int H, int W) {
    __nv_bfloat16* dx;
    cudaMalloc(&dx, (size_t)B * C::CIN * H * W * 2);
    cudaMemset(dx, 0, (size_t)B * C::CIN * H * W * 2);
    CUtensorMap full, tail, map4;
    if (!make_maps<C>(dx, B, H, W, &full, &tail)) return;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C::CIN, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 2, (cuuint64_t)W * H * 2, (cuuint64_t)W * H * C::CIN * 2};
    const cuuint32_t box[4] = {32, (cuuint32_t)C::TH, (cuuint32_t)C::CIN, 1}, estr[4] = {1, 1, 1, 1};
    g_encode(&map4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int fh = H / C::PS, fw = W / C::PS, total = B * fh * fw;
    long long* dcyc;
    cudaMalloc(&dcyc, 1024 * sizeof(long long));
    for (int mode = 0; mode < 2; ++mode)
        for (int ctas = 1; ctas <= 2; ++ctas) {
            const int grid = 148 * ctas, smem = 2 * 64 * 1024 / ctas + 2048 > 100000 ? 110000 : 2 * 64 * 1024 / ctas + 2048;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) { cudaFuncSetAttribute(rate<C, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000); rate<C, 0><<<grid, 32, 110000>>>(full, tail, map4, fh, fw, total, dcyc); }
                else { cudaFuncSetAttribute(rate<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000); rate<C, 1><<<grid, 32, 110000>>>(full, tail, map4, fh, fw, total, dcyc); }
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("  rate: %s\n", cudaGetErrorString(e)); return; }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                best = ms < best ? ms : best;
            }
            (void)smem;
            const double useful = (double)total * C::CIN * C::TH * C::PS * 2;
            printf("  %s rate mode %d (%s) %d CTA/SM: %.1f us for %d patches -> %.0f ns/patch/SM, %.0f GB/s of body bytes\n", name, mode,
                   mode ? "4-D 64-byte rows" : "5-D swizzled rows", ctas, best * 1e3, total, best * 1e6 / (total / 148.0), useful / best * 1e-6);
        }
    cudaFree(dx); cudaFree(dcyc);
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode fn\n"); return 1; }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    int rc = 0;
    rc |= run_case<Cfg<34, 68, 16>>("L4 HyperSeg-M");
    rc |= run_case<Cfg<26, 52, 16>>("L4 S-city");
    rc |= run_case<Cfg<24, 48, 8>>("L3 HyperSeg-M");
    rc |= run_case<Cfg<14, 28, 8>>("L3 S-city");
    run_rate<Cfg<34, 68, 16>>("L4", 8, 256, 512);
    run_rate<Cfg<24, 48, 8>>("L3", 8, 128, 256);
    printf(rc ? "sw32 probe FAILED\n" : "sw32 probe OK\n");
    return rc;
}
