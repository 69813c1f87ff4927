// Fused patch-wise inverted-residual MetaBlock, restage-free tensor-core kernel (tcgen05 / TMEM / TMA), bf16.
//
// Same arithmetic as the reference's HyperPatchInvertedResidual.conv (hyperseg/models/hyperseg_v1_0.py:328-370,
// hyperseg_v1_0_unify.py:342-389): reflect pad 1 -> (ph+2)x(pw+2) halo tile -> W1.tile, BN1, ReLU6 -> depthwise 3x3
// (valid), BN2, ReLU6 -> W3., BN3.  What is different from patch_ir_tc.cu (round 1): no thread ever re-stages an
// operand.
//
//   x tile    The body of the halo tile ((ps+2) rows x ps pixels) is loaded by TMA through a 5-D view of NCHW x,
//             dims {W, c%8, H, c/8, B}, box {ps, 8, ps+2, 1, 1}, SWIZZLE_32B (ps = 16) / no swizzle (ps = 8).  It lands
//             as [c/8][row][c%8][ps px], which IS the canonical MN-major UMMA A operand: M-group = one tile row,
//             K rows 32 (16) bytes apart.  Rows outside the image are TMA zero fill and replaced by their mirror row
//             for border patches.  The 2(ps+2) halo-column pixels are read from global memory one patch ahead
//             (2-byte loads, reflection in the index) and stored into three extra M-groups.
//   weights   The row of a patch arrives in "arranged" order (ir_arranged.cuh): B1 and B2 are K-major UMMA B
//             operands as they lie, W2T is [tap][channel]; BatchNorm scales are already folded in (by the weight
//             head that produced the row or by hsb_ir_arrange_weights).  Two bulk copies per patch.
//   shifts    BN1 / BN3 shifts: one constant "init" MMA per accumulator tile, D = Ones(128x16) . ShiftB, before the
//             data MMAs accumulate on top.  BN2 shift: initial value of the depthwise accumulation.
//   GEMM1     H[(ps+2)^2 px x hid] in TMEM, M = 128 tiles; epilogue 1 = ReLU6 + bf16 -> shared hidden tile.
//   depthwise packed bf16x2 on CUDA cores, thread = 4 channels x 2 adjacent columns x 4 rows, 8-byte shared loads /
//             stores, written straight into GEMM2's 128B-swizzled A operand, half a patch (one M tile) at a time so
//             that GEMM2 of the first half runs under the depthwise of the second.
//   GEMM2     O[ps*ps px x Cout]; epilogue 2 = bf16 -> shared [c][u][v] tile -> ONE TMA store per patch (NCHW).
//
// Persistent CTAs, two (ps = 16) or up to six (ps = 8) per SM; four CTA-wide barriers per patch; x / weight loads
// for patch i+1 are issued as soon as GEMM1 of patch i has retired, so they fly under everything else.
#include <cuda.h>

#include <cstdio>
#include <mutex>

#include "common.cuh"
#include "ir_arranged.cuh"
#include "tcgen05.cuh"

namespace hsb {

void note_kernel(const char* name);

template <int CIN_, int HID_, int COUT_, int PS_>
struct IR2 {
    static constexpr int CIN = CIN_, HID = HID_, COUT = COUT_, PS = PS_, TH = PS_ + 2;
    // arranged weight row
    static constexpr int KC1 = (CIN + 7) / 8, KC2 = (HID + 7) / 8;
    static constexpr int B1_LBO = HID * 16, B2_LBO = COUT * 16;
    static constexpr int SZ_B1 = KC1 * B1_LBO, SZ_W2T = ir_r16(18 * HID), SZ_B2 = KC2 * B2_LBO;
    static constexpr int ROW_BYTES = SZ_B1 + SZ_W2T + SZ_B2;
    // GEMM shapes
    static constexpr int K1 = ir_r16(CIN), N1 = ir_r16(HID), K2 = ir_r16(HID), N2 = ir_r16(COUT);
    // A1 (MN-major): M-group = PS pixels of one channel row group; k-group = 8 channels
    static constexpr int ROWB = PS * 2, GRP = 8 * ROWB;
    static constexpr int HALO = 2 * TH, HALO_G = (HALO + PS - 1) / PS, MG = TH + HALO_G;
    static constexpr int A1_MGS = GRP, A1_KGS = MG * GRP;
    static constexpr int GPT = 128 / PS;                       // M-groups per 128-row tile
    static constexpr int BODY = TH * PS, T = BODY + HALO, M1T = (T + 127) / 128;
    static constexpr int SZ_A1 = ir_r1024(KC1 * A1_KGS);       // the bytes that follow (A2) only ever hold finite bf16
    // A2: one 128-row tile = RPH output rows
    static constexpr int OUT_PX = PS * PS, M2T = (OUT_PX + 127) / 128;
    static constexpr int RPH = PS / M2T;                        // output rows per half
    static constexpr int RQ = RPH / 2;                          // output rows per depthwise thread
    static constexpr int SZ_A2S = 128 * 128, KT2 = K2 > 64 ? K2 - 64 : 0;
    static constexpr int A2T_LBO = 128 * 16, SZ_A2T = (KT2 / 8) * A2T_LBO;
    static constexpr int SZ_A2 = SZ_A2S + SZ_A2T;
    static constexpr int SZ_YST = COUT * OUT_PX * 2;            // output staging tile [c][u][v], inside A2
    // hidden tile [pixel][channel]
    static constexpr int HPITCH = ir_r8(HID) * 2, SZ_HID = ir_r128(T * HPITCH);
    // weight buffers: what the descriptors read past the data must be zero
    static constexpr int SZ_W1 = ir_r128(ir_max(SZ_B1, (K1 / 8 - 1) * B1_LBO + N1 * 16));
    static constexpr int SZ_W23 = ir_r128(SZ_W2T + ir_max(SZ_B2, (K2 / 8 - 1) * B2_LBO + N2 * 16));
    static constexpr int SZ_SHB1 = 2 * N1 * 16, SZ_SHB2 = 2 * N2 * 16;
    static constexpr int MAINQ = ir_min(16, HID / 4), TAILQ = HID / 4 - MAINQ;
    static constexpr int DWW = PS / 2;                          // main depthwise warps: one per column pair
    static constexpr int WARPS = DWW + (TAILQ > 0 ? 1 : 0), THREADS = 32 * WARPS;
    static constexpr int NH = (HALO * CIN + THREADS - 1) / THREADS;   // halo pixels per thread
    static constexpr int TMEM_COLS = ir_pow2_cols(ir_max(M1T * N1, M2T * N2));
    // shared memory map (bytes from a 1024-aligned base)
    static constexpr int OFF_A1 = 0, OFF_A2 = OFF_A1 + SZ_A1, OFF_HID = OFF_A2 + ir_r1024(SZ_A2);
    static constexpr int OFF_W1 = OFF_HID + SZ_HID, OFF_W23 = OFF_W1 + SZ_W1;
    static constexpr int OFF_SHB1 = OFF_W23 + SZ_W23, OFF_SHB2 = OFF_SHB1 + SZ_SHB1, OFF_ONES = OFF_SHB2 + SZ_SHB2;
    static constexpr int OFF_B2B = OFF_ONES + 128, OFF_BAR = OFF_B2B + ir_r16(HID * 2);
    static constexpr int USED_BYTES = OFF_BAR + 128, SMEM_BYTES = USED_BYTES + 1024;
    static constexpr int CTAS = ir_max(1, ir_min(ir_min((228 * 1024 - 1024) / (SMEM_BYTES + 1024), 512 / TMEM_COLS), PS == 16 ? 2 : 6));
    static_assert(PS == 16 || PS == 8, "patch size");
    static_assert(HID % 4 == 0 && TAILQ <= 1, "hidden width: multiple of 4, at most 68");
    static_assert(SZ_YST <= SZ_A2S, "output staging tile must fit the swizzled part of A2");
    static_assert((K1 / 8 - KC1) * A1_KGS + GPT * GRP <= SZ_A2, "GEMM1 over-read must stay inside A2");
    static_assert(3 + M1T + M2T <= 12, "barrier slots");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct IR2Params {
    const __nv_bfloat16* x;
    const __nv_bfloat16* w;      // arranged rows
    const float* shift[3];       // bn1, bn2, bn3 shifts
    int B, H, W, fh, fw, total;
    int64_t w_row_stride;        // elements between rows
#ifdef HSB_IR_PROF
    long long* prof;             // [grid][16] per-phase cycle sums of thread 0 (profiling build only)
#endif
};

#ifdef HSB_IR_PROF
#define HSB_STAMP(k) do { if (tid == 0) { long long now_ = clock64(); prof_acc[k] += now_ - prof_t; prof_t = now_; } } while (0)
#else
#define HSB_STAMP(k) do { } while (0)
#endif

__device__ __forceinline__ uint32_t relu6_pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    const __nv_bfloat162 six = __floats2bfloat162_rn(6.f, 6.f);
    __nv_bfloat162 v = __hmin2(*reinterpret_cast<__nv_bfloat162*>(&r), six);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}

constexpr uint32_t SWZ_32B_MODE = 6;

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::CTAS)
patch_ir2_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap xmap_tail,
                 const __grid_constant__ CUtensorMap ymap, const IR2Params p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint64_t* bar_x = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* bar_w1 = bar_x + 1;
    uint64_t* bar_w23 = bar_x + 2;
    uint64_t* bar_mma1 = bar_x + 3;                       // [M1T]
    uint64_t* bar_mma2 = bar_mma1 + C::M1T;               // [M2T]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_x + 12);
    const uint32_t sm_base = smem_u32(sm);
    const uint32_t a1_addr = sm_base + C::OFF_A1, a2_addr = sm_base + C::OFF_A2, a2t_addr = a2_addr + C::SZ_A2S;
    const uint32_t w1_addr = sm_base + C::OFF_W1, b2_addr = sm_base + C::OFF_W23 + C::SZ_W2T;

    // ---------------- one-time setup ----------------
    for (int i = tid; i < C::OFF_BAR / 16; i += C::THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int n = tid; n < C::HID; n += C::THREADS) {      // ShiftB1: K-major, element (n, k = 0) = BN1 shift
        *reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_SHB1 + n * 16) = __float2bfloat16_rn(p.shift[0][n]);
        reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_B2B)[n] = __float2bfloat16_rn(p.shift[1][n]);
    }
    for (int n = tid; n < C::COUT; n += C::THREADS)
        *reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_SHB2 + n * 16) = __float2bfloat16_rn(p.shift[2][n]);
    for (int i = tid; i < 64; i += C::THREADS) reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_ONES)[i] = __float2bfloat16_rn(1.f);
    if (tid == 0) {
        mbar_init(bar_x, 1);
        mbar_init(bar_w1, 1);
        mbar_init(bar_w23, 1);
        for (int t = 0; t < C::M1T; ++t) mbar_init(bar_mma1 + t, 1);
        for (int t = 0; t < C::M2T; ++t) mbar_init(bar_mma2 + t, 1);
        mbar_fence_init();
        tma_prefetch_desc(&xmap);
        tma_prefetch_desc(&xmap_tail);
        tma_prefetch_desc(&ymap);
    }
    if (warp == 0) tmem_alloc(tmem_slot, C::TMEM_COLS);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int P = p.fh * p.fw;
    const size_t HW = (size_t)p.H * p.W;
    // single-thread roles, elected once
    const bool is_mma = warp == 0 && elect_one();
    const bool is_loader = warp == C::WARPS - 1 && elect_one();
    const bool is_storer = warp == (C::WARPS > 1 ? 1 : 0) && elect_one();

    auto load_x_w1 = [&](int patch) {                     // loader thread: A1 and the B1 buffer are free
        const int b = patch / P, pp = patch % P, pi = pp / p.fw, pj = pp % p.fw;
        mbar_arrive_expect_tx(bar_x, C::KC1 * C::TH * C::GRP);
#pragma unroll
        for (int kg = 0; kg < C::KC1; ++kg) {
            const bool tail = (C::CIN % 8 != 0) && kg == C::KC1 - 1;
            tma_load_5d(sm + C::OFF_A1 + kg * C::A1_KGS, tail ? &xmap_tail : &xmap, pj * C::PS, 0, pi * C::PS - 1, tail ? 0 : kg, b, bar_x);
        }
        mbar_arrive_expect_tx(bar_w1, C::SZ_B1);
        bulk_g2s(sm + C::OFF_W1, p.w + (size_t)patch * p.w_row_stride, C::SZ_B1, bar_w1);
    };
    auto load_w23 = [&](int patch) {                      // loader thread: the W2T / B2 buffer is free
        mbar_arrive_expect_tx(bar_w23, C::SZ_W2T + C::SZ_B2);
        bulk_g2s(sm + C::OFF_W23, p.w + (size_t)patch * p.w_row_stride + C::SZ_B1 / 2, C::SZ_W2T + C::SZ_B2, bar_w23);
    };

    // halo-column pixels: thread-constant task descriptors, the values travel in registers one patch ahead
    uint32_t h_dst[C::NH];      // byte offset in A1 (swizzled), 0xFFFFFFFF = no task
    int h_off[C::NH];           // element offset from the patch's (channel 0, tile row 0, body column 0), interior patches
    uint32_t h_val[C::NH];
#pragma unroll
    for (int j = 0; j < C::NH; ++j) {
        const int i = tid + j * C::THREADS;
        h_dst[j] = 0xFFFFFFFFu;
        h_off[j] = 0;
        h_val[j] = 0;
        if (i < C::HALO * C::CIN) {
            const int c = i / C::HALO, h = i % C::HALO, side = h >= C::TH ? 1 : 0, r = h - side * C::TH;
            uint32_t o = (c >> 3) * C::A1_KGS + (C::TH + h / C::PS) * C::GRP + (c & 7) * C::ROWB + (h % C::PS) * 2;
            if (C::PS == 16) o ^= ((o >> 7) & 1u) << 4;
            h_dst[j] = o;
            h_off[j] = (int)(c * HW) + r * p.W + (side ? C::PS : -1);
        }
    }
    auto prefetch_halo = [&](int patch) {
        const int b = patch / P, pp = patch % P, pi = pp / p.fw, pj = pp % p.fw;
        const int y0 = pi * C::PS - 1, x0 = pj * C::PS;
        const unsigned short* xb = reinterpret_cast<const unsigned short*>(p.x) + (size_t)b * C::CIN * HW;
        const bool border = pi == 0 || pj == 0 || pi == p.fh - 1 || pj == p.fw - 1;
        if (!border) {
            const unsigned short* base = xb + (ptrdiff_t)y0 * p.W + x0;
#pragma unroll
            for (int j = 0; j < C::NH; ++j)
                if (h_dst[j] != 0xFFFFFFFFu) h_val[j] = __ldg(base + h_off[j]);
        } else {
#pragma unroll
            for (int j = 0; j < C::NH; ++j) {
                const int i = tid + j * C::THREADS;
                if (i < C::HALO * C::CIN) {
                    const int c = i / C::HALO, h = i % C::HALO, side = h >= C::TH ? 1 : 0, r = h - side * C::TH;
                    int gy = y0 + r, gx = side ? x0 + C::PS : x0 - 1;
                    gy = gy < 0 ? -gy : (gy >= p.H ? 2 * p.H - 2 - gy : gy);
                    gx = gx < 0 ? -gx : (gx >= p.W ? 2 * p.W - 2 - gx : gx);
                    h_val[j] = __ldg(xb + (size_t)c * HW + (size_t)gy * p.W + gx);
                }
            }
        }
    };
    // registers -> A1 halo groups, mirror rows of border patches; everything the next GEMM1 needs from threads
    auto stage_patch = [&](int patch, uint32_t par) {
#pragma unroll
        for (int j = 0; j < C::NH; ++j)
            if (h_dst[j] != 0xFFFFFFFFu) *reinterpret_cast<unsigned short*>(sm + C::OFF_A1 + h_dst[j]) = (unsigned short)h_val[j];
        const int pp = patch % P, pi = pp / p.fw;
        const bool top = pi == 0, bottom = pi == p.fh - 1;
        if (top || bottom) {                               // tile row 0 <- row 2, row TH-1 <- row TH-3 (whole M-groups)
            mbar_wait(bar_x, par);
            for (int i = tid; i < C::KC1 * (C::GRP / 16); i += C::THREADS) {
                unsigned char* base = sm + C::OFF_A1 + (i / (C::GRP / 16)) * C::A1_KGS + (i % (C::GRP / 16)) * 16;
                if (top) *reinterpret_cast<uint4*>(base) = *reinterpret_cast<const uint4*>(base + 2 * C::GRP);
                if (bottom) *reinterpret_cast<uint4*>(base + (C::TH - 1) * C::GRP) = *reinterpret_cast<const uint4*>(base + (C::TH - 3) * C::GRP);
            }
        }
    };

    constexpr uint32_t IDESC1 = idesc_bf16_f32(128, C::N1, /*A MN-major*/ true, false);
    constexpr uint32_t IDESC1_INIT = idesc_bf16_f32(128, C::N1, false, false);
    constexpr uint32_t IDESC2 = idesc_bf16_f32(128, C::N2, false, false);
    const uint64_t ones_desc = smem_desc(sm_base + C::OFF_ONES, 0, 0, SWZ_NONE);
    // TMEM lanes 32q..32q+31 are reachable only from warps with warp % 4 == q
    const int q = warp & 3, q_warps = (C::WARPS - q + 3) / 4, q_rank = warp >> 2;

    int patch = blockIdx.x;
    if (patch < p.total) {
        if (is_loader) { load_x_w1(patch); load_w23(patch); }
        prefetch_halo(patch);
        stage_patch(patch, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();

#ifdef HSB_IR_PROF
    long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = clock64();
#endif
    uint32_t it = 0;
    for (; patch < p.total; patch += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        HSB_STAMP(11);
        const int next = patch + gridDim.x;
        const int b = patch / P, pp = patch % P, pi = pp / p.fw, pj = pp % p.fw;

        // ---------------- A: GEMM1 (+ BN1 shift), loads and halo prefetch for the next patch ----------------
        if (is_mma) {
            mbar_wait(bar_x, par);
            mbar_wait(bar_w1, par);
            tc_fence_after_sync();
#pragma unroll
            for (int t = 0; t < C::M1T; ++t) {
                umma_bf16(tmem + t * C::N1, ones_desc, smem_desc(sm_base + C::OFF_SHB1, C::N1 * 16, 128, SWZ_NONE), IDESC1_INIT, false);
#pragma unroll
                for (int s = 0; s < C::K1 / 16; ++s) {
                    const uint32_t a = a1_addr + 2 * s * C::A1_KGS + t * C::GPT * C::A1_MGS;
                    const uint64_t da = C::PS == 16 ? smem_desc(a, C::A1_MGS, C::A1_KGS, SWZ_32B_MODE)
                                                    : smem_desc(a, C::A1_KGS, C::A1_MGS, SWZ_NONE);
                    umma_bf16(tmem + t * C::N1, da, smem_desc(w1_addr + 2 * s * C::B1_LBO, C::B1_LBO, 128, SWZ_NONE), IDESC1, true);
                }
                umma_commit(bar_mma1 + t);
            }
        }
        if (next < p.total) {
            if (is_loader) {
                mbar_wait(bar_mma1 + C::M1T - 1, par);     // GEMM1 has retired: A1 and B1 may be overwritten
                load_x_w1(next);
            }
            prefetch_halo(next);
        }

        HSB_STAMP(0);
        // ---------------- B: epilogue 1, tile by tile as the MMAs retire: TMEM -> ReLU6 -> hidden tile ----------------
        for (int t = q_rank; t < C::M1T; t += q_warps) {
            if (t * 128 + q * 32 >= C::T) continue;
            mbar_wait(bar_mma1 + t, par);
            tc_fence_after_sync();
            const int m = t * 128 + q * 32 + lane;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + t * C::N1;
            int hpix;                                      // position of M row m in the TH x TH hidden tile
            if (m < C::BODY) hpix = (m / C::PS) * C::TH + (m % C::PS) + 1;
            else { const int h = m - C::BODY, side = h >= C::TH ? 1 : 0; hpix = (h - side * C::TH) * C::TH + (side ? C::TH - 1 : 0); }
            unsigned char* hrow = sm + C::OFF_HID + (size_t)hpix * C::HPITCH;
            const bool live = m < C::T;
            constexpr int FULL = C::HID / 16, REM = C::HID % 16;
#pragma unroll
            for (int ch = 0; ch < FULL; ch += 2) {
                uint32_t v0[16], v1[16];
                tmem_ld16(taddr + ch * 16, v0);
                if (ch + 1 < FULL) tmem_ld16(taddr + (ch + 1) * 16, v1);
                tmem_ld_wait();
                uint32_t o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = relu6_pack(__uint_as_float(v0[2 * e]), __uint_as_float(v0[2 * e + 1]));
                if (live) {
                    *reinterpret_cast<uint4*>(hrow + ch * 32) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(hrow + ch * 32 + 16) = make_uint4(o[4], o[5], o[6], o[7]);
                }
                if (ch + 1 < FULL) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = relu6_pack(__uint_as_float(v1[2 * e]), __uint_as_float(v1[2 * e + 1]));
                    if (live) {
                        *reinterpret_cast<uint4*>(hrow + (ch + 1) * 32) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4*>(hrow + (ch + 1) * 32 + 16) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            if (REM > 0) {
                static_assert(REM == 0 || REM == 4 || REM == 8 || REM == 12, "hidden width must be a multiple of 4");
                uint32_t v8[8], v4[4];
                if (REM >= 8) tmem_ld8(taddr + FULL * 16, v8);
                constexpr int c4 = FULL * 16 + (REM >= 8 ? 8 : 0);
                if (REM % 8 == 4) tmem_ld4(taddr + c4, v4);
                tmem_ld_wait();
                if (REM >= 8) {
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = relu6_pack(__uint_as_float(v8[2 * e]), __uint_as_float(v8[2 * e + 1]));
                    if (live) *reinterpret_cast<uint4*>(hrow + FULL * 32) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                if (REM % 8 == 4) {
                    const uint32_t o0 = relu6_pack(__uint_as_float(v4[0]), __uint_as_float(v4[1]));
                    const uint32_t o1 = relu6_pack(__uint_as_float(v4[2]), __uint_as_float(v4[3]));
                    if (live) *reinterpret_cast<uint2*>(hrow + c4 * 2) = make_uint2(o0, o1);
                }
            }
        }
        HSB_STAMP(1);
        if (is_storer && it > 0) bulk_wait_read0();        // the previous output tile has left A2
        tc_fence_before_sync();
        __syncthreads();                                   // S2: hidden tile complete, accumulators drained
        HSB_STAMP(2);

        // ---------------- C/D: depthwise 3x3 + BN2 + ReLU6 -> A2, one M tile at a time; GEMM2 behind it ----------------
        {
            const __nv_bfloat162 six = __floats2bfloat162_rn(6.f, 6.f);
            const bool main_thr = warp < C::DWW && (lane & 15) < C::MAINQ;
            const bool tail_thr = C::TAILQ > 0 && warp == C::DWW && (lane >> 3) * 2 < C::RPH;
            // main: quad = lane & 15, columns (2 warp, 2 warp + 1), rows rq*RQ ..;  tail: quad 16, columns (2 (lane & 7), +1), rows 2 (lane >> 3) ..
            const int quad = main_thr ? (lane & 15) : 16;
            const int v0 = main_thr ? 2 * warp : 2 * (lane & 7);
            const int lr0 = main_thr ? (lane >> 4) * C::RQ : (lane >> 3) * 2;     // first output row inside the half
            const int nrows = main_thr ? C::RQ : 2;
            __nv_bfloat162 wt[9][2], bias[2];
            if (main_thr || tail_thr) {
                mbar_wait(bar_w23, par);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const uint2 t2 = *reinterpret_cast<const uint2*>(sm + C::OFF_W23 + k * C::HID * 2 + quad * 8);
                    wt[k][0] = *reinterpret_cast<const __nv_bfloat162*>(&t2.x);
                    wt[k][1] = *reinterpret_cast<const __nv_bfloat162*>(&t2.y);
                }
                const uint2 t2 = *reinterpret_cast<const uint2*>(sm + C::OFF_B2B + quad * 8);
                bias[0] = *reinterpret_cast<const __nv_bfloat162*>(&t2.x);
                bias[1] = *reinterpret_cast<const __nv_bfloat162*>(&t2.y);
            }
#pragma unroll 1
            for (int half = 0; half < C::M2T; ++half) {
                if (main_thr || tail_thr) {
                    const int u0 = half * C::RPH + lr0;                     // first output row = first tile row of the window
                    const unsigned char* src = sm + C::OFF_HID + (size_t)(u0 * C::TH + v0) * C::HPITCH + quad * 8;
                    __nv_bfloat162 r0[4][2], r1[4][2], r2[4][2];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint2 a = *reinterpret_cast<const uint2*>(src + j * C::HPITCH);
                        const uint2 c = *reinterpret_cast<const uint2*>(src + (C::TH + j) * C::HPITCH);
                        r0[j][0] = *reinterpret_cast<const __nv_bfloat162*>(&a.x); r0[j][1] = *reinterpret_cast<const __nv_bfloat162*>(&a.y);
                        r1[j][0] = *reinterpret_cast<const __nv_bfloat162*>(&c.x); r1[j][1] = *reinterpret_cast<const __nv_bfloat162*>(&c.y);
                    }
                    if (half > 0) mbar_wait(bar_mma2 + half - 1, par);     // GEMM2 of the previous half has read A2
                    unsigned char* dst;
                    {
                        const int m = lr0 * C::PS + v0;                     // row of the A2 tile; m & 7 == v0 & 7 for every row step
                        if (main_thr) dst = sm + C::OFF_A2 + m * 128 + ((((quad >> 1) ^ (m & 7)) << 4) | ((quad & 1) << 3));
                        else dst = sm + C::OFF_A2 + C::SZ_A2S + m * 16;
                    }
                    // the second column's row m+1: same 128-byte-row arithmetic with (m + 1) & 7
                    const int m1x = ((lr0 * C::PS + v0 + 1) & 7);
                    const int dst1_delta = main_thr ? 128 + ((((quad >> 1) ^ m1x) << 4) - (((quad >> 1) ^ ((lr0 * C::PS + v0) & 7)) << 4)) : 16;
#pragma unroll
                    for (int u = 0; u < C::RQ; ++u) {
                        if (u < nrows) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint2 a = *reinterpret_cast<const uint2*>(src + ((u + 2) * C::TH + j) * C::HPITCH);
                                r2[j][0] = *reinterpret_cast<const __nv_bfloat162*>(&a.x); r2[j][1] = *reinterpret_cast<const __nv_bfloat162*>(&a.y);
                            }
                            uint32_t outa[2], outb[2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                __nv_bfloat162 pa = bias[e], pb = bias[e];
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx) {
                                    pa = __hfma2(wt[kx][e], r0[kx][e], pa);         pb = __hfma2(wt[kx][e], r0[kx + 1][e], pb);
                                    pa = __hfma2(wt[3 + kx][e], r1[kx][e], pa);     pb = __hfma2(wt[3 + kx][e], r1[kx + 1][e], pb);
                                }
                                pa = __hfma2(wt[6][e], r2[0][e], pa);               pb = __hfma2(wt[6][e], r2[1][e], pb);
                                pa = __hfma2(wt[7][e], r2[1][e], pa);               pb = __hfma2(wt[7][e], r2[2][e], pb);
                                pa = __hfma2_relu(wt[8][e], r2[2][e], pa);          pb = __hfma2_relu(wt[8][e], r2[3][e], pb);
                                pa = __hmin2(pa, six);
                                pb = __hmin2(pb, six);
                                outa[e] = *reinterpret_cast<uint32_t*>(&pa);
                                outb[e] = *reinterpret_cast<uint32_t*>(&pb);
                            }
                            unsigned char* d = dst + u * C::PS * (main_thr ? 128 : 16);
                            *reinterpret_cast<uint2*>(d) = make_uint2(outa[0], outa[1]);
                            *reinterpret_cast<uint2*>(d + dst1_delta) = make_uint2(outb[0], outb[1]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { r0[j][0] = r1[j][0]; r0[j][1] = r1[j][1]; r1[j][0] = r2[j][0]; r1[j][1] = r2[j][1]; }
                        }
                    }
                } else if (half > 0) {
                    mbar_wait(bar_mma2 + half - 1, par);
                }
                HSB_STAMP(3 + 2 * half);
                fence_proxy_async_smem();
                tc_fence_before_sync();
                __syncthreads();                           // S3: this half of A2 is complete
                HSB_STAMP(4 + 2 * half);
                if (is_mma) {
                    tc_fence_after_sync();
                    umma_bf16(tmem + half * C::N2, ones_desc, smem_desc(sm_base + C::OFF_SHB2, C::N2 * 16, 128, SWZ_NONE), IDESC2, false);
#pragma unroll
                    for (int s = 0; s < C::K2 / 16; ++s) {
                        const uint64_t da = s < 4 ? smem_desc(a2_addr + s * 32, 16, 1024, SWZ_128B)
                                                  : smem_desc(a2t_addr + 2 * (s - 4) * C::A2T_LBO, C::A2T_LBO, 128, SWZ_NONE);
                        umma_bf16(tmem + half * C::N2, da, smem_desc(b2_addr + 2 * s * C::B2_LBO, C::B2_LBO, 128, SWZ_NONE), IDESC2, true);
                    }
                    umma_commit(bar_mma2 + half);
                }
            }
        }

        // ---------------- E: epilogue 2: TMEM -> bf16 -> output staging tile [c][u][v] (inside A2) ----------------
        for (int t = q_rank; t < C::M2T; t += q_warps) {
            if (t * 128 + q * 32 >= C::OUT_PX) continue;
            mbar_wait(bar_mma2 + C::M2T - 1, par);           // every GEMM2 tile has retired: A2 is free
            tc_fence_after_sync();
            const int pix = t * 128 + q * 32 + lane;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + t * C::N2;
            unsigned char* ydst = sm + C::OFF_A2 + pix * 2;
#pragma unroll
            for (int c0 = 0; c0 < C::COUT; c0 += 16) {
                uint32_t v[16];
                if (C::COUT - c0 > 8) tmem_ld16(taddr + c0, v);
                else if (C::COUT - c0 > 4) { uint32_t v8[8]; tmem_ld8(taddr + c0, v8);
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = v8[e]; }
                else { uint32_t v4[4]; tmem_ld4(taddr + c0, v4);
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = v4[e]; }
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (c0 + e < C::COUT && pix < C::OUT_PX)
                        *reinterpret_cast<__nv_bfloat16*>(ydst + (c0 + e) * C::OUT_PX * 2) = __float2bfloat16_rn(__uint_as_float(v[e]));
            }
        }

        HSB_STAMP(7);
        // ---------------- F: stage the next patch (halo pixels, mirror rows), request its W2T / B2 ----------------
        if (next < p.total) {
            if (is_loader) {
                mbar_wait(bar_mma2 + C::M2T - 1, par);
                load_w23(next);
            }
            stage_patch(next, par ^ 1);
        }
        HSB_STAMP(8);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();                                   // S41: output tile staged, TMEM drained, A1 halo groups written
        HSB_STAMP(9);
        if (is_storer) {
            tma_store_4d(&ymap, sm + C::OFF_A2, pj * C::PS, pi * C::PS, 0, b);
            bulk_commit();
        }
    }
#ifdef HSB_IR_PROF
    if (tid == 0) {
        for (int k = 0; k < 12; ++k) p.prof[(size_t)blockIdx.x * 16 + k] = prof_acc[k];
        p.prof[(size_t)blockIdx.x * 16 + 12] = it;
    }
#endif
    if (is_storer) bulk_wait0();
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
}

// ---- raw -> arranged rows (for weights that were not produced by the arranged head) -----------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ir_arrange_kernel(const T* __restrict__ w, __nv_bfloat16* __restrict__ out, const float* s1,
                                                         const float* s2, const float* s3, int cin, int hid, int cout, int P,
                                                         int total, WStrides ws, int row_elems) {
    for (int patch = blockIdx.x; patch < total; patch += gridDim.x) {
        const T* src = w + (size_t)(patch / P) * ws.b + (size_t)(patch % P) * ws.p;
        __nv_bfloat16* dst = out + (size_t)patch * row_elems;
        for (int e = threadIdx.x; e < row_elems; e += blockDim.x) {
            const IRSource s = ir_arranged_source(e, cin, hid, cout);
            float v = 0.f;
            if (s.src >= 0) {
                const float sc = s.which == 0 ? s1[s.ch] : (s.which == 1 ? s2[s.ch] : s3[s.ch]);
                v = ld_f(src + (size_t)s.src * ws.k) * sc;
            }
            dst[e] = __float2bfloat16_rn(v);
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 encode_fn() {
    static std::once_flag once;
    static EncodeTiledFn2 fn = nullptr;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn2>(ptr);
        else
            cudaGetLastError();
    });
    return fn;
}

template <class C>
static int launch_ir2(const void* x, void* y, const IR2Params& p, cudaStream_t st) {
    EncodeTiledFn2 encode = encode_fn();
    if (!encode) return fail(HSB_ERR_CUDA, "patch_ir2: cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap xm, xt, ym;
    const cuuint64_t HW2 = (cuuint64_t)p.H * p.W * 2;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const cuuint32_t box[5] = {(cuuint32_t)C::PS, 8, (cuuint32_t)C::TH, 1, 1};
    const CUtensorMapSwizzle swz = C::PS == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const cuuint64_t xstr[4] = {HW2, (cuuint64_t)p.W * 2, 8 * HW2, (cuuint64_t)C::CIN * HW2};
    const cuuint64_t xdim[5] = {(cuuint64_t)p.W, 8, (cuuint64_t)p.H, (cuuint64_t)(C::CIN / 8 > 0 ? C::CIN / 8 : 1), (cuuint64_t)p.B};
    CUresult r = CUDA_SUCCESS;
    if (C::CIN >= 8)
        r = encode(&xm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), xdim, xstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS && C::CIN % 8 != 0) {
        const cuuint64_t tdim[5] = {(cuuint64_t)p.W, (cuuint64_t)(C::CIN % 8), (cuuint64_t)p.H, 1, (cuuint64_t)p.B};
        void* base = const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(x) + (size_t)(C::CIN / 8) * 8 * p.H * p.W);
        r = encode(&xt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, tdim, xstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (C::CIN < 8) xm = xt;
    } else {
        xt = xm;
    }
    if (r == CUDA_SUCCESS) {
        const cuuint64_t ydim[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)C::COUT, (cuuint64_t)p.B};
        const cuuint64_t ystr[3] = {(cuuint64_t)p.W * 2, HW2, (cuuint64_t)C::COUT * HW2};
        const cuuint32_t ybox[4] = {(cuuint32_t)C::PS, (cuuint32_t)C::PS, (cuuint32_t)C::COUT, 1};
        r = encode(&ym, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, ydim, ystr, ybox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(HSB_ERR_CUDA, "patch_ir2: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    auto kern = patch_ir2_kernel<C>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("patch_ir2 attr: ") + cudaGetErrorString(e));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    static const int force_ctas = [] { const char* v = getenv("HSB_IR_CTAS"); return v ? atoi(v) : 0; }();
    const int ctas = force_ctas > 0 ? force_ctas : C::CTAS;
    static const bool verbose = [] { const char* v = getenv("HSB_VERBOSE"); return v && v[0] == '1'; }();
    if (verbose) {
        int resident = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, C::THREADS, (size_t)C::SMEM_BYTES);
        fprintf(stderr, "[hsb] patch_ir2<%d,%d,%d,%d>: %d threads, %d B smem, %d TMEM cols, occupancy %d, launching %d CTAs/SM\n", C::CIN,
                C::HID, C::COUT, C::PS, C::THREADS, C::SMEM_BYTES, C::TMEM_COLS, resident, ctas);
        cudaGetLastError();
    }
    const int grid = std::min(p.total, std::max(1, device_sm_count()) * ctas);
#ifdef HSB_IR_PROF
    {   // profiling build: per-phase cycle sums of thread 0 of every CTA, printed after a synchronising launch
        static long long* dprof = nullptr;
        if (!dprof) cudaMalloc(&dprof, 4096 * 16 * sizeof(long long));
        cudaMemsetAsync(dprof, 0, 4096 * 16 * sizeof(long long), st);
        IR2Params pp = p;
        pp.prof = dprof;
        kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(xm, xt, ym, pp);
        cudaStreamSynchronize(st);
        static long long host[4096 * 16];
        cudaMemcpy(host, dprof, sizeof(long long) * grid * 16, cudaMemcpyDeviceToHost);
        double sum[12] = {0}; double patches = 0;
        for (int g = 0; g < grid; ++g) { for (int k = 0; k < 12; ++k) sum[k] += (double)host[g * 16 + k]; patches += (double)host[g * 16 + 12]; }
        static const char* names[12] = {"A issue GEMM1+prefetch", "B epilogue 1", "S2 barrier", "C depthwise half 0", "S3a barrier", "D depthwise half 1",
                                        "S3b barrier", "E epilogue 2", "F stage next", "S41 barrier", "-", "store issue/loop"};
        double tot = 0; for (int k = 0; k < 12; ++k) tot += sum[k];
        fprintf(stderr, "[hsb-prof] patch_ir2<%d,%d,%d,%d> grid %d, %.0f patches, %.0f cycles/patch (thread 0)\n", C::CIN, C::HID, C::COUT, C::PS,
                grid, patches, tot / patches);
        for (int k = 0; k < 12; ++k) fprintf(stderr, "[hsb-prof]   %-24s %8.0f cycles/patch  %5.1f %%\n", names[k], sum[k] / patches, 100.0 * sum[k] / tot);
        note_kernel("patch_ir2_kernel");
        return check_launch("patch_ir2 launch");
    }
#endif
    kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(xm, xt, ym, p);
    note_kernel("patch_ir2_kernel");
    return check_launch("patch_ir2 launch");
}

#define HSB_IR2_SHAPES(X) \
    X(34, 68, 19, 16)     /* HyperSeg-M level 4 */            \
    X(26, 52, 19, 16)     /* HyperSeg-S Cityscapes level 4 */ \
    X(22, 44, 12, 16)     /* HyperSeg-S CamVid level 4 */     \
    X(24, 48, 16, 8)      /* HyperSeg-M / CamVid level 3 */   \
    X(14, 28, 8, 8)       /* HyperSeg-S Cityscapes level 3 */

}  // namespace hsb

using namespace hsb;

extern "C" int hsb_patch_ir_arranged_supported(int Cin, int hid, int Cout, int ps) {
#define X(CI, HD, CO, PS) if (Cin == CI && hid == HD && Cout == CO && ps == PS) return 1;
    HSB_IR2_SHAPES(X)
#undef X
    return 0;
}

extern "C" int64_t hsb_ir_arranged_row_elems(int Cin, int hid, int Cout) {
    if (Cin <= 0 || hid <= 0 || Cout <= 0) return -1;
    return IRRow(Cin, hid, Cout).bytes / 2;
}

extern "C" int hsb_ir_arrange_weights(const void* w, void* w_arranged, const float* bn1_scale, const float* bn2_scale,
                                      const float* bn3_scale, int B, int Cin, int hid, int Cout, int fh, int fw, int dtype,
                                      int w_layout, int64_t w_row_stride, void* stream) {
    HSB_REQUIRE(w && w_arranged && bn1_scale && bn2_scale && bn3_scale, HSB_ERR_INVALID_ARG, "hsb_ir_arrange_weights: null pointer");
    HSB_REQUIRE(B > 0 && Cin > 0 && hid > 0 && Cout > 0 && fh > 0 && fw > 0, HSB_ERR_INVALID_ARG, "hsb_ir_arrange_weights: bad dimensions");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "hsb_ir_arrange_weights: bad dtype");
    const int64_t hp = (int64_t)Cin * hid + 9 * hid + (int64_t)hid * Cout;
    HSB_REQUIRE(w_layout == HSB_W_NCHW || w_row_stride >= hp, HSB_ERR_INVALID_ARG, "hsb_ir_arrange_weights: row stride < weights per patch");
    const int P = fh * fw, total = B * P, row = IRRow(Cin, hid, Cout).bytes / 2;
    const WStrides ws = make_wstrides(w_layout, hp, P, w_row_stride);
    const int grid = std::min(total, device_sm_count() * 8);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == HSB_F32)
        ir_arrange_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(w), reinterpret_cast<__nv_bfloat16*>(w_arranged),
                                                      bn1_scale, bn2_scale, bn3_scale, Cin, hid, Cout, P, total, ws, row);
    else
        ir_arrange_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<__nv_bfloat16*>(w_arranged),
                                                              bn1_scale, bn2_scale, bn3_scale, Cin, hid, Cout, P, total, ws, row);
    note_kernel("ir_arrange_kernel");
    return check_launch("hsb_ir_arrange_weights");
}

extern "C" int hsb_patch_ir_arranged_fwd(const void* x, const void* w_arranged, void* y, const float* bn1_shift,
                                         const float* bn2_shift, const float* bn3_shift, int B, int Cin, int hid, int Cout,
                                         int H, int W, int fh, int fw, int64_t w_row_stride, void* stream) {
    HSB_REQUIRE(x && w_arranged && y && bn1_shift && bn2_shift && bn3_shift, HSB_ERR_INVALID_ARG, "hsb_patch_ir_arranged_fwd: null pointer");
    HSB_REQUIRE(B > 0 && fh > 0 && fw > 0 && H % fh == 0 && W % fw == 0, HSB_ERR_INVALID_ARG, "hsb_patch_ir_arranged_fwd: bad geometry");
    const int ps = H / fh;
    HSB_REQUIRE(W / fw == ps, HSB_ERR_UNSUPPORTED, "hsb_patch_ir_arranged_fwd: patches must be square");
    HSB_REQUIRE(hsb_patch_ir_arranged_supported(Cin, hid, Cout, ps), HSB_ERR_UNSUPPORTED,
                "hsb_patch_ir_arranged_fwd: no tensor-core instantiation for this (Cin, hid, Cout, patch size)");
    const int64_t row = IRRow(Cin, hid, Cout).bytes / 2;
    HSB_REQUIRE(w_row_stride >= row && (w_row_stride * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(w_arranged) & 15) == 0,
                HSB_ERR_INVALID_ARG, "hsb_patch_ir_arranged_fwd: arranged rows must be 16-byte aligned and at least hsb_ir_arranged_row_elems apart");
    HSB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (W * 2) % 16 == 0 &&
                ((int64_t)H * W * 2) % 16 == 0, HSB_ERR_INVALID_ARG, "hsb_patch_ir_arranged_fwd: x / y must be 16-byte aligned with 16-byte row strides");
    IR2Params p;
    p.x = reinterpret_cast<const __nv_bfloat16*>(x);
    p.w = reinterpret_cast<const __nv_bfloat16*>(w_arranged);
    p.shift[0] = bn1_shift; p.shift[1] = bn2_shift; p.shift[2] = bn3_shift;
    p.B = B; p.H = H; p.W = W; p.fh = fh; p.fw = fw; p.total = B * fh * fw; p.w_row_stride = w_row_stride;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define X(CI, HD, CO, PS) if (Cin == CI && hid == HD && Cout == CO && ps == PS) return launch_ir2<IR2<CI, HD, CO, PS>>(x, y, p, st);
    HSB_IR2_SHAPES(X)
#undef X
    return fail(HSB_ERR_UNSUPPORTED, "hsb_patch_ir_arranged_fwd: unsupported shape");
}
